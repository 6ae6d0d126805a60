// Host-side Fiat-Shamir machinery, kept on the CPU as in the reference:
//   Keccak-f[1600], SHAKE256 (golang.org/x/crypto/sha3, go.mod:9, under
//   common.Rand), STROBE-128 + Merlin v1.0 (github.com/jsign/merlin, go.mod:7)
//   and the reference's wrapper transcript/transcript.go:15-66.
// Points enter the transcript as the 48-byte compressed encodings the GPU
// produces, so no host-side Fp arithmetic exists anywhere in this library.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>

#include "fiber.hpp"
#include "fr.hpp"

namespace cdlh {

inline uint64_t rol64(uint64_t v, int n) { return (v << n) | (v >> (64 - n)); }  // n in 1..63

// one round fully unrolled on 25 local lanes (generated straight-line code: theta, rho+pi, chi, iota)
inline void keccak_f1600_plain(uint64_t a[25]) {
  static const uint64_t RC[24] = {
      0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,
      0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
      0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
      0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,
      0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
      0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
  uint64_t s0 = a[0], s1 = a[1], s2 = a[2], s3 = a[3], s4 = a[4], s5 = a[5], s6 = a[6], s7 = a[7], s8 = a[8], s9 = a[9], s10 = a[10], s11 = a[11], s12 = a[12], s13 = a[13], s14 = a[14], s15 = a[15], s16 = a[16], s17 = a[17], s18 = a[18], s19 = a[19], s20 = a[20], s21 = a[21], s22 = a[22], s23 = a[23], s24 = a[24];
  for (int rnd = 0; rnd < 24; rnd++) {
    const uint64_t c0 = s0 ^ s5 ^ s10 ^ s15 ^ s20;
    const uint64_t c1 = s1 ^ s6 ^ s11 ^ s16 ^ s21;
    const uint64_t c2 = s2 ^ s7 ^ s12 ^ s17 ^ s22;
    const uint64_t c3 = s3 ^ s8 ^ s13 ^ s18 ^ s23;
    const uint64_t c4 = s4 ^ s9 ^ s14 ^ s19 ^ s24;
    const uint64_t d0 = c4 ^ rol64(c1, 1);
    const uint64_t d1 = c0 ^ rol64(c2, 1);
    const uint64_t d2 = c1 ^ rol64(c3, 1);
    const uint64_t d3 = c2 ^ rol64(c4, 1);
    const uint64_t d4 = c3 ^ rol64(c0, 1);
    const uint64_t b0 = s0 ^ d0;
    const uint64_t b16 = rol64(s5 ^ d0, 36);
    const uint64_t b7 = rol64(s10 ^ d0, 3);
    const uint64_t b23 = rol64(s15 ^ d0, 41);
    const uint64_t b14 = rol64(s20 ^ d0, 18);
    const uint64_t b10 = rol64(s1 ^ d1, 1);
    const uint64_t b1 = rol64(s6 ^ d1, 44);
    const uint64_t b17 = rol64(s11 ^ d1, 10);
    const uint64_t b8 = rol64(s16 ^ d1, 45);
    const uint64_t b24 = rol64(s21 ^ d1, 2);
    const uint64_t b20 = rol64(s2 ^ d2, 62);
    const uint64_t b11 = rol64(s7 ^ d2, 6);
    const uint64_t b2 = rol64(s12 ^ d2, 43);
    const uint64_t b18 = rol64(s17 ^ d2, 15);
    const uint64_t b9 = rol64(s22 ^ d2, 61);
    const uint64_t b5 = rol64(s3 ^ d3, 28);
    const uint64_t b21 = rol64(s8 ^ d3, 55);
    const uint64_t b12 = rol64(s13 ^ d3, 25);
    const uint64_t b3 = rol64(s18 ^ d3, 21);
    const uint64_t b19 = rol64(s23 ^ d3, 56);
    const uint64_t b15 = rol64(s4 ^ d4, 27);
    const uint64_t b6 = rol64(s9 ^ d4, 20);
    const uint64_t b22 = rol64(s14 ^ d4, 39);
    const uint64_t b13 = rol64(s19 ^ d4, 8);
    const uint64_t b4 = rol64(s24 ^ d4, 14);
    s0 = b0 ^ (~b1 & b2);
    s1 = b1 ^ (~b2 & b3);
    s2 = b2 ^ (~b3 & b4);
    s3 = b3 ^ (~b4 & b0);
    s4 = b4 ^ (~b0 & b1);
    s5 = b5 ^ (~b6 & b7);
    s6 = b6 ^ (~b7 & b8);
    s7 = b7 ^ (~b8 & b9);
    s8 = b8 ^ (~b9 & b5);
    s9 = b9 ^ (~b5 & b6);
    s10 = b10 ^ (~b11 & b12);
    s11 = b11 ^ (~b12 & b13);
    s12 = b12 ^ (~b13 & b14);
    s13 = b13 ^ (~b14 & b10);
    s14 = b14 ^ (~b10 & b11);
    s15 = b15 ^ (~b16 & b17);
    s16 = b16 ^ (~b17 & b18);
    s17 = b17 ^ (~b18 & b19);
    s18 = b18 ^ (~b19 & b15);
    s19 = b19 ^ (~b15 & b16);
    s20 = b20 ^ (~b21 & b22);
    s21 = b21 ^ (~b22 & b23);
    s22 = b22 ^ (~b23 & b24);
    s23 = b23 ^ (~b24 & b20);
    s24 = b24 ^ (~b20 & b21);
    s0 ^= RC[rnd];
  }
  a[0] = s0; a[1] = s1; a[2] = s2; a[3] = s3; a[4] = s4; a[5] = s5; a[6] = s6; a[7] = s7; a[8] = s8; a[9] = s9; a[10] = s10; a[11] = s11; a[12] = s12; a[13] = s13; a[14] = s14; a[15] = s15; a[16] = s16; a[17] = s17; a[18] = s18; a[19] = s19; a[20] = s20; a[21] = s21; a[22] = s22; a[23] = s23; a[24] = s24;
}

// Every sponge goes through here: inside a fiber group the state is parked and permuted together
// with the other proofs' states by one eight-way AVX-512 call (fiber.hpp); otherwise in place.
inline void keccak_f1600(uint64_t a[25]) {
  if (!fiber_keccak(a)) keccak_f1600_plain(a);
}

// dst[0..n) ^= src[0..n), eight bytes at a time
inline void xor_bytes(uint8_t* dst, const uint8_t* src, size_t n) {
  size_t i = 0;
  for (; i + 8 <= n; i += 8) {
    uint64_t a, b;
    memcpy(&a, dst + i, 8);
    memcpy(&b, src + i, 8);
    a ^= b;
    memcpy(dst + i, &a, 8);
  }
  for (; i < n; i++) dst[i] ^= src[i];
}

// SHAKE256 with an incremental squeeze (sha3.NewShake256: Write then Read)
class Shake256 {
 public:
  Shake256() { memset(st_, 0, sizeof st_); }
  void absorb(const uint8_t* data, size_t n) {
    uint8_t* s = reinterpret_cast<uint8_t*>(st_);
    while (n) {
      size_t take = n < kRate - pos_ ? n : kRate - pos_;
      xor_bytes(s + pos_, data, take);
      pos_ += take;
      data += take;
      n -= take;
      if (pos_ == kRate) { keccak_f1600(st_); pos_ = 0; }
    }
  }
  void read(uint8_t* out, size_t n) {
    uint8_t* s = reinterpret_cast<uint8_t*>(st_);
    if (!squeezing_) {
      s[pos_] ^= 0x1f;
      s[kRate - 1] ^= 0x80;
      keccak_f1600(st_);
      pos_ = 0;
      squeezing_ = true;
    }
    while (n) {
      if (pos_ == kRate) { keccak_f1600(st_); pos_ = 0; }
      size_t take = n < kRate - pos_ ? n : kRate - pos_;
      memcpy(out, s + pos_, take);
      pos_ += take;
      out += take;
      n -= take;
    }
  }

 private:
  static constexpr size_t kRate = 136;
  uint64_t st_[25];
  size_t pos_ = 0;
  bool squeezing_ = false;
};

// common.Rand (common/rand.go): deterministic SHAKE256 stream + rejection sampling
class Rand {
 public:
  explicit Rand(uint64_t seed) {  // rand.go:19-33
    uint8_t b[8];
    for (int i = 0; i < 8; i++) b[i] = (uint8_t)(seed >> (56 - 8 * i));
    shake_.absorb(b, 8);
  }
  Fr get_fr() {  // rand.go:35-47
    for (;;) {
      uint8_t buf[32];
      shake_.read(buf, 32);
      Fr r;
      if (fr_from_bytes_be_canonical(r, buf)) return r;
    }
  }
  void get_frs(Fr* out, size_t n) {  // rand.go:49-59
    for (size_t i = 0; i < n; i++) out[i] = get_fr();
  }
  std::vector<uint32_t> generate_permutation(size_t n) {  // rand.go:97-113
    std::vector<uint32_t> perm(n);
    for (size_t i = 0; i < n; i++) perm[i] = (uint32_t)i;
    for (size_t i = 0; i < n; i++) {
      uint8_t tmp[16];
      shake_.read(tmp, 16);
      uint32_t v = ((uint32_t)tmp[0] << 8) | tmp[1];
      size_t j = v % (i + 1);
      std::swap(perm[i], perm[j]);
    }
    return perm;
  }

 private:
  Shake256 shake_;
};

// STROBE-128 as instantiated by Merlin
class Strobe128 {
 public:
  explicit Strobe128(const char* proto_label) {
    memset(st_, 0, sizeof st_);
    uint8_t* s = bytes();
    const uint8_t init[6] = {1, kR + 2, 1, 0, 1, 96};
    memcpy(s, init, 6);
    memcpy(s + 6, "STROBEv1.0.2", 12);
    keccak_f1600(st_);
    meta_ad(reinterpret_cast<const uint8_t*>(proto_label), strlen(proto_label), false);
  }
  void meta_ad(const uint8_t* d, size_t n, bool more) { begin_op(kM | kA, more); absorb(d, n); }
  void ad(const uint8_t* d, size_t n, bool more) { begin_op(kA, more); absorb(d, n); }
  void prf(uint8_t* out, size_t n, bool more) { begin_op(kI | kA | kC, more); squeeze(out, n); }

 private:
  static constexpr uint8_t kR = 166;
  static constexpr uint8_t kI = 1, kA = 2, kC = 4, kT = 8, kM = 16, kK = 32;
  uint8_t* bytes() { return reinterpret_cast<uint8_t*>(st_); }
  void run_f() {
    uint8_t* s = bytes();
    s[pos_] ^= pos_begin_;
    s[pos_ + 1] ^= 0x04;
    s[kR + 1] ^= 0x80;
    keccak_f1600(st_);
    pos_ = 0;
    pos_begin_ = 0;
  }
  void absorb(const uint8_t* d, size_t n) {
    uint8_t* s = bytes();
    while (n) {
      size_t take = n < (size_t)(kR - pos_) ? n : (size_t)(kR - pos_);
      xor_bytes(s + pos_, d, take);
      pos_ = (uint8_t)(pos_ + take);
      d += take;
      n -= take;
      if (pos_ == kR) run_f();
    }
  }
  void squeeze(uint8_t* out, size_t n) {
    uint8_t* s = bytes();
    while (n) {
      size_t take = n < (size_t)(kR - pos_) ? n : (size_t)(kR - pos_);
      memcpy(out, s + pos_, take);
      memset(s + pos_, 0, take);
      pos_ = (uint8_t)(pos_ + take);
      out += take;
      n -= take;
      if (pos_ == kR) run_f();
    }
  }
  void begin_op(uint8_t flags, bool more) {
    if (more) return;  // continuation keeps the current flags
    uint8_t old_begin = pos_begin_;
    pos_begin_ = (uint8_t)(pos_ + 1);
    cur_flags_ = flags;
    uint8_t hdr[2] = {old_begin, flags};
    absorb(hdr, 2);
    if ((flags & (kC | kK)) && pos_ != 0) run_f();
  }
  uint64_t st_[25];
  uint8_t pos_ = 0, pos_begin_ = 0, cur_flags_ = 0;
};

// transcript.Transcript over merlin.Transcript
class Transcript {
 public:
  explicit Transcript(const char* label) : strobe_("Merlin v1.0") {  // transcript.go:15-19
    append_message("dom-sep", reinterpret_cast<const uint8_t*>(label), strlen(label));
  }
  void append_message(const char* label, const uint8_t* msg, size_t n) {
    strobe_.meta_ad(reinterpret_cast<const uint8_t*>(label), strlen(label), false);
    uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
    strobe_.meta_ad(len, 4, true);
    strobe_.ad(msg, n, false);
  }
  // AppendPoints / AppendPointsAffine: one message per 48-byte compressed point (transcript.go:25-39)
  void append_points(const char* label, const uint8_t* enc48, size_t count) {
    for (size_t i = 0; i < count; i++) append_message(label, enc48 + 48 * i, 48);
  }
  void append_scalar(const char* label, const Fr& s) {  // transcript.go:41-46
    uint8_t b[32];
    fr_to_bytes_be(b, s);
    append_message(label, b, 32);
  }
  void append_scalars(const char* label, const Fr* s, size_t n) {
    for (size_t i = 0; i < n; i++) append_scalar(label, s[i]);
  }
  // merlin Transcript.ChallengeBytes
  void challenge_bytes(const char* label, uint8_t* dest, size_t n) {
    strobe_.meta_ad(reinterpret_cast<const uint8_t*>(label), strlen(label), false);
    uint8_t len[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
    strobe_.meta_ad(len, 4, true);
    strobe_.prf(dest, n, false);
  }
  Fr challenge(const char* label) {  // GetAndAppendChallenge, transcript.go:48-58
    for (;;) {
      uint8_t dest[32];
      challenge_bytes(label, dest, 32);
      Fr c;
      if (fr_from_bytes_be_canonical(c, dest)) {
        // AppendScalars(label, challenge): Bytes() of a canonical value is the very string it was set from
        append_message(label, dest, 32);
        return c;
      }
    }
  }

 private:
  Strobe128 strobe_;
};

}  // namespace cdlh
