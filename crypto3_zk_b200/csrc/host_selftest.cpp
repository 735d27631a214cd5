// Host-side self-test shim: compiles the SAME arithmetic headers the kernels use with g++
// (PTX primitives emulated) and exposes them over a tiny C ABI so tests/test_host_arith.py can
// compare against Python big integers.  Test vehicle only - not part of libzkb200.so.
#include <stdint.h>
#include <string.h>
#include "zkb_field.cuh"
#include "zkb_curve.cuh"

using namespace zkb;

template <class P>
static void field_op(int op, const uint32_t *a, const uint32_t *b, uint32_t *out) {
    Fp<P> x, y, r;
    memcpy(x.l, a, sizeof(x.l));
    memcpy(y.l, b, sizeof(y.l));
    switch (op) {
        case 0: r = x * y; break;            // montgomery product
        case 1: r = x + y; break;
        case 2: r = x - y; break;
        case 3: r = x.neg(); break;
        case 4: r = x.to_mont(); break;
        case 5: r = x.from_mont(); break;
        case 6: r = x.inverse(); break;      // montgomery in/out
        case 7: r = (x.to_mont() * y.to_mont()).from_mont(); break;  // canonical product
        default: r = Fp<P>::zero();
    }
    memcpy(out, r.l, sizeof(r.l));
}

template <class P>
static void curve_op(int op, const uint32_t *a, const uint32_t *b, uint32_t *out) {
    // a, b: affine points (x,y) canonical, all-zero = infinity; out: affine canonical
    typedef Fp<P> F;
    Affine<F> pa, pb;
    memcpy(pa.x.l, a, sizeof(F)); memcpy(pa.y.l, a + P::N, sizeof(F));
    memcpy(pb.x.l, b, sizeof(F)); memcpy(pb.y.l, b + P::N, sizeof(F));
    pa = pa.to_mont(); pb = pb.to_mont();
    XYZZ<F> acc = XYZZ<F>::infinity();
    switch (op) {
        case 0: acc.add_mixed(pa); acc.add_mixed(pb); break;              // a + b (mixed)
        case 1: { XYZZ<F> q = XYZZ<F>::infinity(); q.add_mixed(pb); acc.add_mixed(pa); acc.add(q); } break;
        case 2: acc.add_mixed(pa); acc = acc.dbl(); break;               // 2a
        case 3: { XYZZ<F> q = XYZZ<F>::infinity(); q.add_mixed(pa); q = q.dbl(); q.add_mixed(pb);
                  acc.add_mixed(pa); acc.add(q); } break;                 // a + (2a + b)
    }
    Affine<F> r = acc.to_affine().from_mont();
    memcpy(out, r.x.l, sizeof(F)); memcpy(out + P::N, r.y.l, sizeof(F));
}

extern "C" int zkb_host_field_op(int field_id, int op, const uint32_t *a, const uint32_t *b, uint32_t *out) {
    switch (field_id) {
        case 0: field_op<params::Bls12381Fr>(op, a, b, out); return 0;
        case 1: field_op<params::Bn254Fr>(op, a, b, out); return 0;
        case 2: field_op<params::PallasFp>(op, a, b, out); return 0;
        case 3: field_op<params::PallasFq>(op, a, b, out); return 0;
        case 4: field_op<params::Bls12381Fq>(op, a, b, out); return 0;
        case 5: field_op<params::Bn254Fq>(op, a, b, out); return 0;
    }
    return 1;
}

extern "C" int zkb_host_curve_op(int curve_id, int op, const uint32_t *a, const uint32_t *b, uint32_t *out) {
    switch (curve_id) {
        case 0: curve_op<params::Bls12381Fq>(op, a, b, out); return 0;
        case 1: curve_op<params::Bn254Fq>(op, a, b, out); return 0;
        case 2: curve_op<params::PallasFp>(op, a, b, out); return 0;
    }
    return 1;
}

// ------------------------------------------------------------------------------------------------
// CPU replay of the NTT pass kernels: same plan, same tables, same per-thread phase functions as the
// device path, executed thread by thread.  Validates the index logic without a GPU.
#include <vector>
#include "zkb_ntt_plan.h"
#include "zkb_ntt_tables.cuh"

template <class P>
static void replay_pass(const NttPassParams &p) {
    typedef Fp<P> F;
    std::vector<u128> smem(ntt_smem_bytes(p.log_r) / 16);
    const uint32_t nt = ZKB_NTT_THREADS;
    uint64_t tiles = ntt_pass_tiles(p);
    for (uint64_t tile = 0; tile < tiles; tile++) {
        const NttTile t = ntt_tile(p, tile);
#define H_ALL(stmt) for (uint32_t tid = 0; tid < nt; tid++) { stmt; }
#define H_L H_ALL(if (p.tw_in_smem) ntt_phase_stage_tw<F>(p, smem.data(), tid, nt); ntt_phase_load<F>(p, t, smem.data(), tid, nt))
#define H_PM H_ALL(ntt_phase_premul<F>(p, t, smem.data(), tid, nt))
#define H_R2(lh) H_ALL(ntt_phase_radix2<F>(p, smem.data(), lh, tid, nt))
#define H_R4(lh) H_ALL(ntt_phase_radix4<F>(p, smem.data(), lh, tid, nt))
#define H_S H_ALL(ntt_phase_store<F>(p, t, smem.data(), tid, nt))
        ZKB_NTT_FOR_EACH_PHASE(H_L, H_PM, H_R2, H_R4, H_S, p.log_r, p.zero_levels, p.load_tab != nullptr)
    }
}

template <class P>
static int host_ntt(int log_n, int log_n_in, int inverse, const uint32_t *shift, uint32_t batch,
                    const uint32_t *in, uint32_t *out, const uint32_t *known_src = nullptr, int known_log = 0) {
    typedef Fp<P> F;
    if (log_n > P::TWO_ADICITY || log_n < 1) return 2;
    const uint64_t N = 1ull << log_n, Nin = 1ull << log_n_in;
    NttPlan pl = ntt_make_plan(log_n, Nin < N);   // as ntt_device_t chooses
    if (pl.n_passes < 1) pl = ntt_make_plan(log_n);
    if (pl.n_passes < 1) return 3;
    F wN = ntt_omega<F, P>(log_n, inverse);
    F wT = ntt_omega<F, P>(ZKB_NTT_TW_LOG, inverse);
    F ninv = ntt_n_inv<F>(log_n);
    std::vector<F> tw(1u << (ZKB_NTT_TW_LOG - 1));
    for (size_t i = 0; i < tw.size(); i++) tw[i] = powtab_value<F>(wT, F::one(), 0, 0, i);
    std::vector<std::vector<F>> inter(ZKB_NTT_MAX_PASSES);
    NttTables tb = {};
    tb.tw = tw.data();
    tb.tw_in_smem = 1;
    for (int i = 0; i + 1 < pl.n_passes; i++) {
        int lmp = pl.log_m(i), lm = pl.log_m(i + 1);
        F base = wN;
        for (int k = 0; k < pl.log_n - lmp; k++) base = base.sqr();
        F scale = (i == 0 && inverse) ? ninv : F::one();
        inter[i].resize(1ull << lmp);
        for (uint64_t idx = 0; idx < (1ull << lmp); idx++) inter[i][idx] = powtab_value<F>(base, scale, 1, lm, idx);
        tb.inter[i] = inter[i].data();
    }
    std::vector<F> cos;
    if (shift) {
        F g;
        memcpy(g.l, shift, sizeof(g.l));
        g = g.to_mont();
        if (!inverse) {   // a[i] *= g^i before the transform
            cos.resize(N);
            for (uint64_t i = 0; i < N; i++) cos[i] = powtab_value<F>(g, F::one(), 0, 0, i);
            tb.load_tab = cos.data(); tb.load_mask = N - 1;
        } else {          // a[i] *= g^-i after the transform (and 1/N when it is not in T_1)
            F gi = g.inverse();
            cos.resize(N);
            for (uint64_t i = 0; i < N; i++)
                cos[i] = powtab_value<F>(gi, pl.n_passes == 1 ? ninv : F::one(), 0, 0, i);
            tb.store_tab = cos.data(); tb.store_mask = N - 1;
        }
    } else if (inverse && pl.n_passes == 1) {
        cos.resize(1);
        cos[0] = ninv;
        tb.store_tab = cos.data(); tb.store_mask = 0;
    }
    std::vector<F> work((size_t)N * batch);
    // poison the work buffer: a pass that reads something the known-output path never stored would show up
    memset(work.data(), 0xA5, work.size() * sizeof(F));
    auto passes = ntt_build_passes(pl, tb, in, out, work.data(), batch, Nin, N, Nin, known_src, known_log, Nin);
    for (auto &q : passes) replay_pass<P>(q);
    if (passes.back().known_log > 0) {
        const int kl = passes.back().known_log;
        const uint64_t total = (uint64_t)batch << (log_n - kl);
        for (uint64_t e = 0; e < total; e++) {
            uint64_t s_, d_;
            ntt_known_scatter_index(e, log_n, kl, Nin, N, &s_, &d_);
            memcpy(out + 8 * d_, known_src + 8 * s_, 32);
        }
    }
    return 0;
}

// polynomial_dfs::resize as lde_device runs it: inverse transform, then the zero-padded forward transform whose
// outputs at multiples of the blow-up are taken from the input evaluations
template <class P>
static int host_lde(int log_in, int log_out, uint32_t batch, const uint32_t *in, uint32_t *out) {
    std::vector<uint32_t> coef((size_t)batch << log_in << 3);
    int rc = host_ntt<P>(log_in, log_in, 1, nullptr, batch, in, coef.data());
    if (rc) return rc;
    return host_ntt<P>(log_out, log_in, 0, nullptr, batch, coef.data(), out, in, log_out - log_in);
}

extern "C" int zkb_host_lde(int field_id, int log_in, int log_out, uint32_t batch, const uint32_t *in, uint32_t *out) {
    switch (field_id) {
        case 0: return host_lde<params::Bls12381Fr>(log_in, log_out, batch, in, out);
        case 1: return host_lde<params::Bn254Fr>(log_in, log_out, batch, in, out);
        case 2: return host_lde<params::PallasFp>(log_in, log_out, batch, in, out);
        case 3: return host_lde<params::PallasFq>(log_in, log_out, batch, in, out);
    }
    return 1;
}

extern "C" int zkb_host_ntt(int field_id, int log_n, int log_n_in, int inverse, const uint32_t *shift,
                            uint32_t batch, const uint32_t *in, uint32_t *out) {
    switch (field_id) {
        case 0: return host_ntt<params::Bls12381Fr>(log_n, log_n_in, inverse, shift, batch, in, out);
        case 1: return host_ntt<params::Bn254Fr>(log_n, log_n_in, inverse, shift, batch, in, out);
        case 2: return host_ntt<params::PallasFp>(log_n, log_n_in, inverse, shift, batch, in, out);
        case 3: return host_ntt<params::PallasFq>(log_n, log_n_in, inverse, shift, batch, in, out);
    }
    return 1;
}
