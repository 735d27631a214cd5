// Carry-chain primitives for multiprecision arithmetic on 32-bit limbs.
//
// Device (sm_100a): thin wrappers over add.cc / addc / mad.lo.cc / madc.hi.cc PTX; ptxas turns the
// lo/hi pairs into IMAD.WIDE.U32 + carry-predicate chains on the integer (fma) pipe.
// Host (g++): the same primitives emulated with an explicit thread-local carry flag, so that the
// exact algorithms in zkb_field.cuh / zkb_curve.cuh can be unit-tested in the GPU-less build
// container (tests/test_host_arith.py).  The host path is a test vehicle for the device code and
// is also used for the handful of host-side scalar operations (twiddle seeds, final MSM window
// combine) - it is never a fallback for a kernel.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ZKB_HD __host__ __device__ __forceinline__
#define ZKB_D __device__ __forceinline__
#define ZKB_HD_NOINLINE __host__ __device__ __noinline__
#else
#define ZKB_HD inline
#define ZKB_D inline
#define ZKB_HD_NOINLINE inline
#endif

namespace zkb {
namespace ptx {

#if defined(__CUDA_ARCH__)

ZKB_D uint32_t add_cc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
ZKB_D uint32_t addc_cc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
ZKB_D uint32_t addc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
ZKB_D uint32_t sub_cc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
ZKB_D uint32_t subc_cc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
ZKB_D uint32_t subc(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
ZKB_D uint32_t mul_lo(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
ZKB_D uint32_t mul_hi(uint32_t a, uint32_t b) {
    uint32_t r; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
}
ZKB_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
ZKB_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
ZKB_D uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
ZKB_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
ZKB_D uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r; asm volatile("madc.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}
ZKB_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
}

// 64-bit accumulator words (an aligned register pair each): mul.wide + add[c].cc.u64 is the PTX shape
// ptxas turns into ONE IMAD.WIDE.U32[.X] with carry-in/out predicates, whatever the other operand is
// (register or immediate), and pack() of two 32-bit halves keeps shifts/adds on the ALU pipe.
ZKB_D uint64_t mul_wide(uint32_t a, uint32_t b) {
    uint64_t r; asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(r) : "r"(a), "r"(b)); return r;
}
ZKB_D uint64_t mad_wide_cc(uint32_t a, uint32_t b, uint64_t c) {
    uint64_t r; asm volatile("{.reg .u64 t; mul.wide.u32 t, %1, %2; add.cc.u64 %0, t, %3;}" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r;
}
ZKB_D uint64_t madc_wide_cc(uint32_t a, uint32_t b, uint64_t c) {
    uint64_t r; asm volatile("{.reg .u64 t; mul.wide.u32 t, %1, %2; addc.cc.u64 %0, t, %3;}" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r;
}
ZKB_D uint64_t madc_wide(uint32_t a, uint32_t b, uint64_t c) {
    uint64_t r; asm volatile("{.reg .u64 t; mul.wide.u32 t, %1, %2; addc.u64 %0, t, %3;}" : "=l"(r) : "r"(a), "r"(b), "l"(c)); return r;
}
ZKB_D uint64_t add_cc64(uint64_t a, uint64_t b) {
    uint64_t r; asm volatile("add.cc.u64 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
ZKB_D uint64_t addc_cc64(uint64_t a, uint64_t b) {
    uint64_t r; asm volatile("addc.cc.u64 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
ZKB_D uint64_t pack64(uint32_t lo, uint32_t hi) {
    uint64_t r; asm volatile("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi)); return r;
}

#else  // ---------------------------------------------------------------- host emulation

inline uint32_t &cc_flag() { static thread_local uint32_t cc = 0; return cc; }

inline uint32_t add_cc(uint32_t a, uint32_t b) {
    uint64_t t = (uint64_t)a + b; cc_flag() = (uint32_t)(t >> 32); return (uint32_t)t;
}
inline uint32_t addc_cc(uint32_t a, uint32_t b) {
    uint64_t t = (uint64_t)a + b + cc_flag(); cc_flag() = (uint32_t)(t >> 32); return (uint32_t)t;
}
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + cc_flag(); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) {
    uint64_t t = (uint64_t)a - b; cc_flag() = (uint32_t)((t >> 32) & 1); return (uint32_t)t;
}
inline uint32_t subc_cc(uint32_t a, uint32_t b) {
    uint64_t t = (uint64_t)a - b - cc_flag(); cc_flag() = (uint32_t)((t >> 32) & 1); return (uint32_t)t;
}
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - cc_flag(); }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return (uint32_t)((uint64_t)a * b); }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_lo(a, b), c); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_lo(a, b), c); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_hi(a, b), c); }
inline uint32_t madc_lo(uint32_t a, uint32_t b, uint32_t c) { return addc(mul_lo(a, b), c); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return addc(mul_hi(a, b), c); }

inline uint64_t mul_wide(uint32_t a, uint32_t b) { return (uint64_t)a * b; }
inline uint64_t add_cc64(uint64_t a, uint64_t b) {
    unsigned __int128 t = (unsigned __int128)a + b; cc_flag() = (uint32_t)(t >> 64); return (uint64_t)t;
}
inline uint64_t addc_cc64(uint64_t a, uint64_t b) {
    unsigned __int128 t = (unsigned __int128)a + b + cc_flag(); cc_flag() = (uint32_t)(t >> 64); return (uint64_t)t;
}
inline uint64_t mad_wide_cc(uint32_t a, uint32_t b, uint64_t c) { return add_cc64(mul_wide(a, b), c); }
inline uint64_t madc_wide_cc(uint32_t a, uint32_t b, uint64_t c) { return addc_cc64(mul_wide(a, b), c); }
inline uint64_t madc_wide(uint32_t a, uint32_t b, uint64_t c) { return mul_wide(a, b) + c + cc_flag(); }
inline uint64_t pack64(uint32_t lo, uint32_t hi) { return (uint64_t)lo | ((uint64_t)hi << 32); }

#endif

}  // namespace ptx
}  // namespace zkb
