// Host tails of the MSM (compiled with g++): Horner combine of the per-window sums
// sum_w 2^(c w) S_w and the final XYZZ -> affine conversion.  O(c W) group operations in total;
// everything proportional to the number of points runs in the kernels of zkb_msm.cu.
#include <string.h>
#include "../../include/zkb200.h"
#include "zkb_curve.cuh"
#include "zkb_hostfield.h"

using namespace zkb;

template <class P>
static XYZZ<HostFp<P>> load_xyzz(const uint32_t *l) {
    XYZZ<HostFp<P>> r;
    r.X = HostFp<P>::from_limbs32(l);
    r.Y = HostFp<P>::from_limbs32(l + P::N);
    r.ZZ = HostFp<P>::from_limbs32(l + 2 * P::N);
    r.ZZZ = HostFp<P>::from_limbs32(l + 3 * P::N);
    return r;
}
template <class P>
static void store_xyzz(const XYZZ<HostFp<P>> &p, uint32_t *l) {
    p.X.to_limbs32(l);
    p.Y.to_limbs32(l + P::N);
    p.ZZ.to_limbs32(l + 2 * P::N);
    p.ZZZ.to_limbs32(l + 3 * P::N);
}

template <class P>
static void window_combine(int c, int W, const uint32_t *sums, uint32_t *out) {
    XYZZ<HostFp<P>> acc = XYZZ<HostFp<P>>::infinity();
    for (int w = W - 1; w >= 0; w--) {
        if (w != W - 1)
            for (int i = 0; i < c; i++) acc = acc.dbl();
        acc.add(load_xyzz<P>(sums + (size_t)w * 4 * P::N));
    }
    store_xyzz<P>(acc, out);
}

template <class P>
static void combine(uint32_t count, const uint32_t *partials, uint32_t *result_affine) {
    XYZZ<HostFp<P>> acc = XYZZ<HostFp<P>>::infinity();
    for (uint32_t i = 0; i < count; i++) acc.add(load_xyzz<P>(partials + (size_t)i * 4 * P::N));
    Affine<HostFp<P>> a = acc.to_affine().from_mont();
    a.x.to_limbs32(result_affine);
    a.y.to_limbs32(result_affine + P::N);
}

namespace zkb {
// sums: W window sums (XYZZ, Montgomery, 4*N limbs each) -> one XYZZ partial result
int msm_window_combine(int curve, int c, int W, const uint32_t *sums, uint32_t *out) {
    switch (curve) {
        case ZKB_CURVE_BLS12_381_G1: window_combine<params::Bls12381Fq>(c, W, sums, out); return ZKB_OK;
        case ZKB_CURVE_BN254_G1: window_combine<params::Bn254Fq>(c, W, sums, out); return ZKB_OK;
        case ZKB_CURVE_PALLAS: window_combine<params::PallasFp>(c, W, sums, out); return ZKB_OK;
    }
    return ZKB_ERR_INVALID_ARGUMENT;
}
}  // namespace zkb

extern "C" int zkb_msm_combine(int curve, uint32_t count, const uint32_t *partials_xyzz, uint32_t *result_affine) {
    if (!partials_xyzz || !result_affine) return ZKB_ERR_INVALID_ARGUMENT;
    switch (curve) {
        case ZKB_CURVE_BLS12_381_G1: combine<params::Bls12381Fq>(count, partials_xyzz, result_affine); return ZKB_OK;
        case ZKB_CURVE_BN254_G1: combine<params::Bn254Fq>(count, partials_xyzz, result_affine); return ZKB_OK;
        case ZKB_CURVE_PALLAS: combine<params::PallasFp>(count, partials_xyzz, result_affine); return ZKB_OK;
    }
    return ZKB_ERR_INVALID_ARGUMENT;
}

template <class G, int N>
static void write_gen(uint32_t *out) {
    for (int i = 0; i < N; i++) { out[i] = G::x(i); out[N + i] = G::y(i); }
}
// curve_type::g1_type<>::value_type::one() of the reference (affine x || y, canonical limbs)
extern "C" int zkb_curve_generator(int curve, uint32_t *out_affine) {
    if (!out_affine) return ZKB_ERR_INVALID_ARGUMENT;
    switch (curve) {
        case ZKB_CURVE_BLS12_381_G1: write_gen<params::GenBls12381G1, 12>(out_affine); return ZKB_OK;
        case ZKB_CURVE_BN254_G1: write_gen<params::GenBn254G1, 8>(out_affine); return ZKB_OK;
        case ZKB_CURVE_PALLAS: write_gen<params::GenPallas, 8>(out_affine); return ZKB_OK;
    }
    return ZKB_ERR_INVALID_ARGUMENT;
}
