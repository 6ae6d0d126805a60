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

#include "fr.hpp"

namespace cdlh {

inline uint64_t rol64(uint64_t v, int n) { return n ? (v << n) | (v >> (64 - n)) : v; }

inline void keccak_f1600(uint64_t a[25]) {
  static const uint64_t RC[24] = {
      0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull,
      0x000000000000808bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
      0x000000000000008aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000aull,
      0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull, 0x8000000000008003ull,
      0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
      0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
  static const int ROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
  for (int rnd = 0; rnd < 24; rnd++) {
    uint64_t c[5], d[5], b[25];
    for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
    for (int x = 0; x < 5; x++) d[x] = c[(x + 4) % 5] ^ rol64(c[(x + 1) % 5], 1);
    for (int i = 0; i < 25; i++) a[i] ^= d[i % 5];
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) b[y + 5 * ((2 * x + 3 * y) % 5)] = rol64(a[x + 5 * y], ROT[x + 5 * y]);
    for (int x = 0; x < 5; x++)
      for (int y = 0; y < 5; y++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
    a[0] ^= RC[rnd];
  }
}

// SHAKE256 with an incremental squeeze (sha3.NewShake256: Write then Read)
class Shake256 {
 public:
  Shake256() { memset(st_, 0, sizeof st_); }
  void absorb(const uint8_t* data, size_t n) {
    uint8_t* s = reinterpret_cast<uint8_t*>(st_);
    for (size_t i = 0; i < n; i++) {
      s[pos_++] ^= data[i];
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
    for (size_t i = 0; i < n; i++) {
      if (pos_ == kRate) { keccak_f1600(st_); pos_ = 0; }
      out[i] = s[pos_++];
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
    for (size_t i = 0; i < n; i++) {
      s[pos_++] ^= d[i];
      if (pos_ == kR) run_f();
    }
  }
  void squeeze(uint8_t* out, size_t n) {
    uint8_t* s = bytes();
    for (size_t i = 0; i < n; i++) {
      out[i] = s[pos_];
      s[pos_++] = 0;
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
        append_scalar(label, c);
        return c;
      }
    }
  }

 private:
  Strobe128 strobe_;
};

}  // namespace cdlh
