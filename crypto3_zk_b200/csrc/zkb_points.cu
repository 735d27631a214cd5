// Batched decompression of curve points in the reference's wire encoding: a Groth16 proving key
// blob holds its query vectors as compressed BLS12-381 points, and the reader takes one square root per point
// (Fq for G1, Fq2 for G2) - a 381-bit exponentiation each.  One thread per point on the MSM's field code.
//
// Reference: zk/snark/systems/ppzksnark/r1cs_gg_ppzksnark/marshalling.hpp:97-198 (g1 / g2 / sparse_vector /
// knowledge-commitment readers call curve_element_serializer<bls12<381>>::octets_to_g1_point / octets_to_g2_point once per
// element, :656-738 the proving key).  The serializer itself lives in crypto3-algebra (not vendored); it is the ZCash
// encoding: x big-endian (G2: x.c1 then x.c0), bit 7 of byte 0 = compressed, bit 6 = infinity, bit 5 = y is the
// lexicographically larger of (y, -y) (Fq2: decided by c1, by c0 when c1 = 0).
#include <stdio.h>
#include <string.h>
#define ZKB_MUL_OUTLINE 1
#include "zkb_curve.cuh"
#include "zkb_internal.h"

using namespace zkb;

typedef Fp<params::Bls12381Fq> FqBls;
typedef Fp2<FqBls> Fq2Bls;

enum { PT_OK = 0, PT_INFINITY = 1, PT_NOT_COMPRESSED = 2, PT_BAD_INFINITY = 3, PT_NOT_REDUCED = 4, PT_NOT_ON_CURVE = 5 };

struct SqrtExponents {
    uint32_t p_plus_1_div_4[12];    // Fq square root (p = 3 mod 4)
    uint32_t p_minus_3_div_4[12];   // Fq2 square root, a1 = a^((p-3)/4)
    uint32_t p_minus_1_div_2[12];   //                  b = (1 + alpha)^((p-1)/2)
};

// 48 big-endian bytes (the three flag bits of byte 0 cleared) -> 12 little-endian limbs
__device__ __forceinline__ FqBls load_be48(const uint8_t *b) {
    FqBls v;
#pragma unroll
    for (int k = 0; k < 12; k++) {
        const uint8_t *q = b + 44 - 4 * k;
        uint32_t w = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
        if (k == 11) w &= 0x1fffffffu;
        v.l[k] = w;
    }
    return v;
}
__device__ __forceinline__ bool fq_reduced(const FqBls &v) {
    for (int i = 11; i >= 0; i--) {
        if (v.l[i] < params::Bls12381Fq::mod(i)) return true;
        if (v.l[i] > params::Bls12381Fq::mod(i)) return false;
    }
    return false;
}
// canonical v > p - v, i.e. v > (p - 1) / 2 (false for 0)
__device__ __forceinline__ bool fq_sign(const FqBls &canonical) {
    if (canonical.is_zero()) return false;
    const FqBls n = canonical.neg();       // p - v on canonical limbs (neg does not depend on the form)
    for (int i = 11; i >= 0; i--) {
        if (canonical.l[i] > n.l[i]) return true;
        if (canonical.l[i] < n.l[i]) return false;
    }
    return false;
}
template <class T>
__device__ __noinline__ T pow_limbs12(const T &a, const uint32_t *e) {
    T r = T::one();
    bool started = false;
    for (int i = 11; i >= 0; i--) {
        for (int b = 31; b >= 0; b--) {
            if (started) r = r.sqr();
            if ((e[i] >> b) & 1) {
                r = started ? r * a : a;
                started = true;
            }
        }
    }
    return r;
}

__global__ void __launch_bounds__(128) g1_decompress_kernel(uint64_t n, const uint8_t *__restrict__ in, uint64_t stride, SqrtExponents ex,
                                                            Affine<FqBls> *__restrict__ out, uint8_t *__restrict__ status,
                                                            uint32_t *__restrict__ first_bad) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *b = in + i * stride;
    const uint8_t flags = b[0];
    Affine<FqBls> pt = Affine<FqBls>::infinity();
    int st = PT_OK;
    const FqBls x = load_be48(b);
    if (!(flags & 0x80)) {
        st = PT_NOT_COMPRESSED;
    } else if (flags & 0x40) {
        st = (x.is_zero() && !(flags & 0x20)) ? PT_INFINITY : PT_BAD_INFINITY;
    } else if (!fq_reduced(x)) {
        st = PT_NOT_REDUCED;
    } else {
        const FqBls xm = x.to_mont();
        const FqBls rhs = xm.sqr() * xm + FqBls::one().dbl().dbl();      // x^3 + 4
        FqBls y = pow_limbs12(rhs, ex.p_plus_1_div_4);
        if (y.sqr() != rhs) {
            st = PT_NOT_ON_CURVE;
        } else {
            y = y.from_mont();
            if (fq_sign(y) != ((flags & 0x20) != 0)) y = y.neg();
            pt.x = x;
            pt.y = y;
        }
    }
    out[i] = pt;
    if (status) status[i] = (uint8_t)st;
    if (st > PT_INFINITY) atomicMin(first_bad, (uint32_t)(i < 0xfffffffeull ? i : 0xfffffffeull));
}

__global__ void __launch_bounds__(128) g2_decompress_kernel(uint64_t n, const uint8_t *__restrict__ in, uint64_t stride, SqrtExponents ex,
                                                            Affine<Fq2Bls> *__restrict__ out, uint8_t *__restrict__ status,
                                                            uint32_t *__restrict__ first_bad) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint8_t *b = in + i * stride;
    const uint8_t flags = b[0];
    Affine<Fq2Bls> pt = Affine<Fq2Bls>::infinity();
    int st = PT_OK;
    Fq2Bls x;
    x.c1 = load_be48(b);
    x.c0 = load_be48(b + 48);
    {   // the second half has no flag bits: restore what load_be48 masked
        const uint32_t top = ((uint32_t)b[48] << 24) | ((uint32_t)b[49] << 16) | ((uint32_t)b[50] << 8) | b[51];
        x.c0.l[11] = top;
    }
    if (!(flags & 0x80)) {
        st = PT_NOT_COMPRESSED;
    } else if (flags & 0x40) {
        st = (x.is_zero() && !(flags & 0x20)) ? PT_INFINITY : PT_BAD_INFINITY;
    } else if (!fq_reduced(x.c0) || !fq_reduced(x.c1)) {
        st = PT_NOT_REDUCED;
    } else {
        const Fq2Bls xm = x.to_mont();
        const FqBls four = FqBls::one().dbl().dbl();
        Fq2Bls rhs = xm.sqr() * xm;                                     // x^3 + 4 (1 + u)
        rhs.c0 = rhs.c0 + four;
        rhs.c1 = rhs.c1 + four;
        Fq2Bls y = Fq2Bls::zero();
        bool ok = true;
        if (!rhs.is_zero()) {
            // square root in Fq2 for p = 3 mod 4: a1 = a^((p-3)/4), alpha = a1^2 a, x0 = a1 a;
            // alpha = -1: y = u x0, else y = (1 + alpha)^((p-1)/2) x0; checked by squaring
            const Fq2Bls a1 = pow_limbs12(rhs, ex.p_minus_3_div_4);
            const Fq2Bls alpha = a1.sqr() * rhs;
            const Fq2Bls x0 = a1 * rhs;
            if (alpha.c1.is_zero() && alpha.c0 == FqBls::one().neg()) {
                y.c0 = x0.c1.neg();
                y.c1 = x0.c0;
            } else {
                Fq2Bls t = alpha;
                t.c0 = t.c0 + FqBls::one();
                y = pow_limbs12(t, ex.p_minus_1_div_2) * x0;
            }
            ok = y.sqr() == rhs;
        }
        if (!ok) {
            st = PT_NOT_ON_CURVE;
        } else {
            y = y.from_mont();
            const bool sign = y.c1.is_zero() ? fq_sign(y.c0) : fq_sign(y.c1);
            if (sign != ((flags & 0x20) != 0)) y = y.neg();
            pt.x = x;
            pt.y = y;
        }
    }
    out[i] = pt;
    if (status) status[i] = (uint8_t)st;
    if (st > PT_INFINITY) atomicMin(first_bad, (uint32_t)(i < 0xfffffffeull ? i : 0xfffffffeull));
}

// (p + add) >> shift on the 12 limbs of the modulus, add in {-3, -1, +1}: the lowest limb of p is 0xffffaaab, so the
// addition neither carries nor borrows
static void modulus_shifted(int add, int shift, uint32_t *out) {
    uint32_t t[12];
    for (int i = 0; i < 12; i++) t[i] = params::Bls12381Fq::mod(i);
    t[0] = (uint32_t)((int64_t)t[0] + add);
    for (int i = 0; i < 12; i++) out[i] = (t[i] >> shift) | (i + 1 < 12 ? t[i + 1] << (32 - shift) : 0);
}

extern "C" int zkb_points_decompress(zkb_ctx *ctx, int curve, uint64_t n, const uint8_t *octets, uint64_t stride_bytes,
                                     void *points_affine_out, uint8_t *status_out, int mem, void *stream) {
    if (!ctx) return ZKB_ERR_INVALID_ARGUMENT;
    const bool g2 = curve == ZKB_CURVE_BLS12_381_G2;
    if (curve != ZKB_CURVE_BLS12_381_G1 && !g2)
        return ctx_fail(ctx, ZKB_ERR_UNSUPPORTED, "zkb_points_decompress: the reference's serializer covers BLS12-381 G1 / G2");
    const uint64_t width = g2 ? 96 : 48;
    if (n == 0) return ZKB_OK;
    if (!octets || !points_affine_out || stride_bytes < width || n >= (1ull << 32))
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_points_decompress: bad arguments");
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t in_bytes = (size_t)(n - 1) * stride_bytes + width, out_bytes = (size_t)n * 2 * (g2 ? 24 : 12) * 4;
    const uint8_t *d_in = octets;
    void *d_out = points_affine_out;
    uint8_t *d_status = status_out;
    if (mem != ZKB_MEM_DEVICE) {
        void *p;
        ZKB_TRY(ctx_scratch(ctx, "pts_in", in_bytes, &p));
        ZKB_CUDA_OK(ctx, cudaMemcpyAsync(p, octets, in_bytes, cudaMemcpyHostToDevice, st));
        d_in = (const uint8_t *)p;
        ZKB_TRY(ctx_scratch(ctx, "pts_out", out_bytes, &d_out));
        if (status_out) {
            ZKB_TRY(ctx_scratch(ctx, "pts_status", n, &p));
            d_status = (uint8_t *)p;
        }
    }
    void *flag;
    ZKB_TRY(ctx_scratch(ctx, "pts_flag", 16, &flag));
    ZKB_CUDA_OK(ctx, cudaMemsetAsync(flag, 0xff, 4, st));
    SqrtExponents ex;
    modulus_shifted(1, 2, ex.p_plus_1_div_4);
    modulus_shifted(-3, 2, ex.p_minus_3_div_4);
    modulus_shifted(-1, 1, ex.p_minus_1_div_2);
    const unsigned blocks = (unsigned)((n + 127) / 128);
    if (g2)
        g2_decompress_kernel<<<blocks, 128, 0, st>>>(n, d_in, stride_bytes, ex, (Affine<Fq2Bls> *)d_out, d_status, (uint32_t *)flag);
    else
        g1_decompress_kernel<<<blocks, 128, 0, st>>>(n, d_in, stride_bytes, ex, (Affine<FqBls> *)d_out, d_status, (uint32_t *)flag);
    ctx->launches++;
    ZKB_CUDA_OK(ctx, cudaGetLastError());
    uint32_t bad = 0;
    ZKB_CUDA_OK(ctx, cudaMemcpyAsync(&bad, flag, 4, cudaMemcpyDeviceToHost, st));
    if (mem != ZKB_MEM_DEVICE) {
        ZKB_CUDA_OK(ctx, cudaMemcpyAsync(points_affine_out, d_out, out_bytes, cudaMemcpyDeviceToHost, st));
        if (status_out) ZKB_CUDA_OK(ctx, cudaMemcpyAsync(status_out, d_status, n, cudaMemcpyDeviceToHost, st));
    }
    ZKB_CUDA_OK(ctx, cudaStreamSynchronize(st));
    if (bad != 0xffffffffu) {
        char msg[160];
        snprintf(msg, sizeof msg, "zkb_points_decompress: point %u is not a valid compressed point (status_out tells why)", bad);
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, msg);
    }
    return ZKB_OK;
}
