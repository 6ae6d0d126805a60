// Carry-chain primitives for the 32-bit-limb field cores.
//
// On the device every helper is exactly one PTX instruction of the
// add.cc/addc/mad.lo.cc/madc.hi.cc family (ptxas pairs mad.lo.cc+madc.hi.cc on
// the same operands into one IMAD.WIDE.U32[.X] with a predicate carry, which is
// what makes a 12-limb Montgomery product cost 2*12^2+12 integer-pipe
// multiply-adds).  On the host the same helpers are emulated bit-exactly with an
// explicit carry flag, so the field/curve code above them is unit-tested on the
// CPU build (tests/ "hostcheck") before it ever reaches a GPU.
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define CDL_HD __host__ __device__ __forceinline__
#define CDL_D __device__ __forceinline__
// Group-law / inversion routines are force-inlined on the device.  (Real calls
// were tried: ptxas passes their pointer arguments in per-warp uniform
// registers, which a lane that early-returns and runs ahead clobbers for its
// warp-mates — an illegal-address fault on sm_100a with CUDA 12.9.)  Kernels
// keep the number of inlined copies low with non-unrolled loops instead.
#define CDL_FN static __host__ __device__ __forceinline__
#else
#define CDL_HD inline
#define CDL_D inline
#define CDL_FN static inline
#endif

namespace cdl {

// Carry context: empty on device (hardware CC.CF), explicit flag on host.
struct CC {
#if !defined(__CUDA_ARCH__)
  uint32_t cf = 0;
#endif
};

#if defined(__CUDA_ARCH__)

CDL_HD uint32_t add_cc(CC&, uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
CDL_HD uint32_t addc_cc(CC&, uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
CDL_HD uint32_t addc(CC&, uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
CDL_HD uint32_t sub_cc(CC&, uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
CDL_HD uint32_t subc_cc(CC&, uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
CDL_HD uint32_t subc(CC&, uint32_t a, uint32_t b) {
  uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
CDL_HD uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
CDL_HD uint32_t mul_hi(uint32_t a, uint32_t b) { return __umulhi(a, b); }
CDL_HD uint32_t mad_lo_cc(CC&, uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
CDL_HD uint32_t madc_lo_cc(CC&, uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
CDL_HD uint32_t mad_hi_cc(CC&, uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
CDL_HD uint32_t madc_hi_cc(CC&, uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
CDL_HD uint32_t madc_hi(CC&, uint32_t a, uint32_t b, uint32_t c) {
  uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}

#else  // host emulation, bit-exact

CDL_HD uint32_t add_cc(CC& cc, uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a + b; cc.cf = (uint32_t)(t >> 32); return (uint32_t)t;
}
CDL_HD uint32_t addc_cc(CC& cc, uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a + b + cc.cf; cc.cf = (uint32_t)(t >> 32); return (uint32_t)t;
}
CDL_HD uint32_t addc(CC& cc, uint32_t a, uint32_t b) { return a + b + cc.cf; }
CDL_HD uint32_t sub_cc(CC& cc, uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a - b; cc.cf = (uint32_t)(t >> 63); return (uint32_t)t;
}
CDL_HD uint32_t subc_cc(CC& cc, uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a - b - cc.cf; cc.cf = (uint32_t)(t >> 63); return (uint32_t)t;
}
CDL_HD uint32_t subc(CC& cc, uint32_t a, uint32_t b) { return a - b - cc.cf; }
CDL_HD uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
CDL_HD uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
CDL_HD uint32_t mad_lo_cc(CC& cc, uint32_t a, uint32_t b, uint32_t c) {
  uint64_t t = (uint64_t)(uint32_t)(a * b) + c; cc.cf = (uint32_t)(t >> 32); return (uint32_t)t;
}
CDL_HD uint32_t madc_lo_cc(CC& cc, uint32_t a, uint32_t b, uint32_t c) {
  uint64_t t = (uint64_t)(uint32_t)(a * b) + c + cc.cf; cc.cf = (uint32_t)(t >> 32); return (uint32_t)t;
}
CDL_HD uint32_t mad_hi_cc(CC& cc, uint32_t a, uint32_t b, uint32_t c) {
  uint64_t t = (((uint64_t)a * b) >> 32) + c; cc.cf = (uint32_t)(t >> 32); return (uint32_t)t;
}
CDL_HD uint32_t madc_hi_cc(CC& cc, uint32_t a, uint32_t b, uint32_t c) {
  uint64_t t = (((uint64_t)a * b) >> 32) + c + cc.cf; cc.cf = (uint32_t)(t >> 32); return (uint32_t)t;
}
CDL_HD uint32_t madc_hi(CC& cc, uint32_t a, uint32_t b, uint32_t c) {
  return (uint32_t)((((uint64_t)a * b) >> 32) + c + cc.cf);
}

#endif

}  // namespace cdl
