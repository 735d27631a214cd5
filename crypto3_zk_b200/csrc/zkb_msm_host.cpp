// Host tails of the MSM (compiled with g++): Horner combine of the per-window sums
// sum_w 2^(c w) S_w and the final XYZZ -> affine conversion.  O(c W) group operations in total;
// everything proportional to the number of points runs in the kernels of zkb_msm.cu.
#include <string.h>
#include "../../include/zkb200.h"
#include "zkb_curve.cuh"
#include "zkb_hostfield.h"

using namespace zkb;

// limb (de)serialisation of host field elements: HostFp<P> (N limbs) and Fp2<HostFp<P>> (c0 || c1)
template <class P>
static void load_f(HostFp<P> &f, const uint32_t *l) { f = HostFp<P>::from_limbs32(l); }
template <class P>
static void store_f(const HostFp<P> &f, uint32_t *l) { f.to_limbs32(l); }
template <class B>
static void load_f(Fp2<B> &f, const uint32_t *l) { load_f(f.c0, l); load_f(f.c1, l + B::N); }
template <class B>
static void store_f(const Fp2<B> &f, uint32_t *l) { store_f(f.c0, l); store_f(f.c1, l + B::N); }

template <class F>
static XYZZ<F> load_xyzz(const uint32_t *l) {
    XYZZ<F> r;
    load_f(r.X, l);
    load_f(r.Y, l + F::N);
    load_f(r.ZZ, l + 2 * F::N);
    load_f(r.ZZZ, l + 3 * F::N);
    return r;
}
template <class F>
static void store_xyzz(const XYZZ<F> &p, uint32_t *l) {
    store_f(p.X, l);
    store_f(p.Y, l + F::N);
    store_f(p.ZZ, l + 2 * F::N);
    store_f(p.ZZZ, l + 3 * F::N);
}

// Horner over the digit windows: window w has base + (w < rem) bits (MsmWindows in zkb_msm.cu)
template <class F>
static void window_combine(int W, int base, int rem, const uint32_t *sums, uint32_t *out) {
    XYZZ<F> acc = XYZZ<F>::infinity();
    for (int w = W - 1; w >= 0; w--) {
        if (w != W - 1) {
            const int width = base + (w < rem ? 1 : 0);
            for (int i = 0; i < width; i++) acc = acc.dbl();
        }
        acc.add(load_xyzz<F>(sums + (size_t)w * 4 * F::N));
    }
    store_xyzz<F>(acc, out);
}

template <class F>
static void combine(uint32_t count, const uint32_t *partials, uint32_t *result_affine) {
    XYZZ<F> acc = XYZZ<F>::infinity();
    for (uint32_t i = 0; i < count; i++) acc.add(load_xyzz<F>(partials + (size_t)i * 4 * F::N));
    Affine<F> a = acc.to_affine().from_mont();
    store_f(a.x, result_affine);
    store_f(a.y, result_affine + F::N);
}

typedef HostFp<params::Bls12381Fq> HFqBls;
typedef HostFp<params::Bn254Fq> HFqBn;
typedef HostFp<params::PallasFp> HFpPallas;
#define ZKB_HOST_DISPATCH_CURVE(curve, ...)                                        \
    switch (curve) {                                                               \
        case ZKB_CURVE_BLS12_381_G1: { typedef HFqBls CF; __VA_ARGS__; } break;       \
        case ZKB_CURVE_BN254_G1: { typedef HFqBn CF; __VA_ARGS__; } break;            \
        case ZKB_CURVE_PALLAS: { typedef HFpPallas CF; __VA_ARGS__; } break;          \
        case ZKB_CURVE_BLS12_381_G2: { typedef Fp2<HFqBls> CF; __VA_ARGS__; } break;  \
        case ZKB_CURVE_BN254_G2: { typedef Fp2<HFqBn> CF; __VA_ARGS__; } break;       \
        default: break;                                                            \
    }

namespace zkb {
// sums: W window sums (XYZZ, Montgomery, 4 * coordinate limbs each) -> one XYZZ partial result
int msm_window_combine(int curve, int W, int base, int rem, const uint32_t *sums, uint32_t *out) {
    ZKB_HOST_DISPATCH_CURVE(curve, window_combine<CF>(W, base, rem, sums, out); return ZKB_OK)
    return ZKB_ERR_INVALID_ARGUMENT;
}
}  // namespace zkb

extern "C" int zkb_msm_combine(int curve, uint32_t count, const uint32_t *partials_xyzz, uint32_t *result_affine) {
    if (!partials_xyzz || !result_affine) return ZKB_ERR_INVALID_ARGUMENT;
    ZKB_HOST_DISPATCH_CURVE(curve, combine<CF>(count, partials_xyzz, result_affine); return ZKB_OK)
    return ZKB_ERR_INVALID_ARGUMENT;
}

template <class G, int N>
static void write_gen(uint32_t *out) {
    for (int i = 0; i < N; i++) { out[i] = G::x(i); out[N + i] = G::y(i); }
}
// curve_type::g1_type<>::value_type::one() / g2_type<>::value_type::one() of the reference (affine x || y, canonical limbs)
extern "C" int zkb_curve_generator(int curve, uint32_t *out_affine) {
    if (!out_affine) return ZKB_ERR_INVALID_ARGUMENT;
    switch (curve) {
        case ZKB_CURVE_BLS12_381_G1: write_gen<params::GenBls12381G1, 12>(out_affine); return ZKB_OK;
        case ZKB_CURVE_BN254_G1: write_gen<params::GenBn254G1, 8>(out_affine); return ZKB_OK;
        case ZKB_CURVE_PALLAS: write_gen<params::GenPallas, 8>(out_affine); return ZKB_OK;
        case ZKB_CURVE_BLS12_381_G2: write_gen<params::GenBls12381G2, 24>(out_affine); return ZKB_OK;
        case ZKB_CURVE_BN254_G2: write_gen<params::GenBn254G2, 16>(out_affine); return ZKB_OK;
    }
    return ZKB_ERR_INVALID_ARGUMENT;
}
