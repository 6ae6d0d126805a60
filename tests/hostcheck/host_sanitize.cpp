// Test-only harness: the product's host-side Fiat-Shamir code (csrc/host/: Keccak / STROBE / Merlin, common.Rand,
// Fr arithmetic, the thread pool and the eight-way fiber hashing) built WITHOUT nvcc so that it can run under the
// compiler's sanitizers (tests/test_host_sanitizers.py builds it with -fsanitize=undefined, =address, =thread).
// It repeats what cdl_host_selftest / cdl_host_selftest_fibers do inside the library and prints "ok".
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#include "../../go-curdleproofs_b200/csrc/host/fiber.hpp"
#include "../../go-curdleproofs_b200/csrc/host/fr.hpp"
#include "../../go-curdleproofs_b200/csrc/host/pool.hpp"
#include "../../go-curdleproofs_b200/csrc/host/transcript.hpp"

using cdlh::Fr;

static int fail(const char* what) {
  fprintf(stderr, "host_sanitize: %s\n", what);
  return 1;
}

int main(int argc, char** argv) {
  const uint32_t n = argc > 1 ? (uint32_t)atoi(argv[1]) : 37, rounds = argc > 2 ? (uint32_t)atoi(argv[2]) : 12;

  // Merlin's published "test protocol" vector (merlin/src/transcript.rs, equivalence_simple)
  {
    static const uint8_t want[32] = {0xd5, 0xa2, 0x19, 0x72, 0xd0, 0xd5, 0xfe, 0x32, 0x0c, 0x0d, 0x26, 0x3f, 0xac, 0x7f, 0xff, 0xb8,
                                     0x14, 0x5a, 0xa6, 0x40, 0xaf, 0x6e, 0x9b, 0xca, 0x17, 0x7c, 0x03, 0xc7, 0xef, 0xcf, 0x06, 0x15};
    cdlh::Transcript t("test protocol");
    t.append_message("some label", reinterpret_cast<const uint8_t*>("some data"), 9);
    uint8_t out[32];
    t.challenge_bytes("challenge", out, 32);
    if (memcmp(out, want, 32)) return fail("Merlin known answer");
  }

  // Fr identities on the deterministic stream of common.Rand
  {
    cdlh::Rand rnd(7);
    std::vector<Fr> v(64);
    rnd.get_frs(v.data(), v.size());
    v[5] = cdlh::FR_ZERO;
    std::vector<Fr> bi = cdlh::fr_batch_inv(v);
    for (size_t i = 0; i < v.size(); i++) {
      const Fr &a = v[i], &b = v[(i + 1) % v.size()];
      if (!cdlh::fr_eq(cdlh::fr_sub(cdlh::fr_add(a, b), b), a)) return fail("fr add / sub");
      if (!cdlh::fr_eq(cdlh::fr_add(a, cdlh::fr_neg(a)), cdlh::FR_ZERO)) return fail("fr neg");
      if (!cdlh::fr_eq(cdlh::fr_mul(a, b), cdlh::fr_mul(b, a))) return fail("fr mul commutes");
      if (!cdlh::fr_eq(cdlh::fr_pow_u64(a, 3), cdlh::fr_mul(cdlh::fr_sqr(a), a))) return fail("fr pow");
      if (cdlh::fr_is_zero(a)) {
        if (!cdlh::fr_is_zero(bi[i])) return fail("batch inversion keeps zeros");
        continue;
      }
      if (!cdlh::fr_eq(cdlh::fr_mul(a, cdlh::fr_inv(a)), cdlh::FR_ONE)) return fail("fr inv");
      if (!cdlh::fr_eq(bi[i], cdlh::fr_inv(a))) return fail("fr batch inv");
      uint8_t be[32];
      cdlh::fr_to_bytes_be(be, a);
      Fr back;
      if (!cdlh::fr_from_bytes_be_canonical(back, be) || !cdlh::fr_eq(back, a)) return fail("fr bytes round trip");
    }
    Fr ip = cdlh::fr_inner(v.data(), bi.data(), v.size());  // 63 non-zero entries times their inverses
    if (!cdlh::fr_eq(ip, cdlh::fr_from_u64(63))) return fail("fr inner product");
    std::vector<uint32_t> perm = rnd.generate_permutation(124);
    std::vector<uint8_t> seen(124, 0);
    for (uint32_t p : perm) {
      if (p >= 124 || seen[p]) return fail("permutation");
      seen[p] = 1;
    }
  }

  // per-proof transcripts: plain, then on a thread pool with eight proofs per thread as fibers
  auto work = [&](size_t i, uint8_t* out) {
    cdlh::Transcript t("curdleproofs");
    cdlh::Rand rnd(1000 + i);
    uint8_t msg[200];
    for (uint32_t r = 0; r < rounds; r++) {
      size_t len = 1 + (i * 7 + r * 13) % sizeof msg;
      for (size_t k = 0; k < len; k++) msg[k] = (uint8_t)(i + 31 * r + k);
      t.append_message("selftest_msg", msg, len);
      Fr c = t.challenge("selftest_challenge");
      Fr x = rnd.get_fr();
      t.append_scalar("selftest_rand", x);
      cdlh::fr_to_bytes_be(out + 32 * r, cdlh::fr_add(c, x));
    }
    t.challenge_bytes("selftest_final", out + 32 * rounds, 32);
  };
  const size_t per = (size_t)rounds * 32 + 32;
  std::vector<uint8_t> plain(n * per), fib(n * per);
  for (size_t i = 0; i < n; i++) work(i, plain.data() + i * per);
  cdlh::ThreadPool pool(3);
  std::function<void(size_t)> fn([&](size_t i) { work(i, fib.data() + i * per); });
  for (int rep = 0; rep < 3; rep++) {  // the pool is persistent: several generations of parallel_for
    std::fill(fib.begin(), fib.end(), 0);
    pool.parallel_for((n + 7) / 8, std::function<void(size_t)>([&](size_t g) {
      cdlh::run_fiber_group(fn, 8 * g, n - 8 * g < 8 ? n - 8 * g : 8);
    }));
    if (plain != fib) return fail("fiber / pool transcripts differ from the plain ones");
  }
  printf("ok fibers=%d n=%u rounds=%u\n", cdlh::fibers_available() ? 1 : 0, n, rounds);
  return 0;
}
