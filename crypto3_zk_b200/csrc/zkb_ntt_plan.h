// Host-side planning for the multi-pass NTT: radix split, per-pass parameter blocks and the
// definition of every twiddle table.  Pure C++ (no CUDA) so that the runtime (zkb_ntt.cu) and the
// CPU replay harness (host_selftest.cpp) share it.
//
// Decomposition (N = R_1 R_2 .. R_p, M_i = R_{i+1} .. R_p, M_0 = N):
//   input index  n = n_1 M_1 + n_2 M_2 + .. + n_p          (n_1 most significant)
//   output index k = k_1 + R_1 (k_2 + R_2 (k_3 + ..))      (k_1 least significant)
//   pass i < p : for all other digits, R_i-point DFT over n_i (stride M_i), then multiply by
//                T_i[k_i M_i + m] = w_{M_{i-1}}^(k_i m),  m = index mod M_i          (in place)
//   pass p     : R_p-point DFT over n_p (stride 1), scatter to natural order        (out of place)
// Inverse: w -> w^-1 everywhere and 1/N folded into T_1 (or a scalar store table when p == 1).
#pragma once
#include <stdint.h>
#include <vector>
#include "zkb_ntt_pass.cuh"

namespace zkb {

struct NttPlan {
    int log_n;
    int n_passes;
    int lr[ZKB_NTT_MAX_PASSES];
    int log_m(int i) const {  // log2 M_i, i in [0, n_passes]
        int s = 0;
        for (int j = i; j < n_passes; j++) s += lr[j];
        return s;
    }
};

// small_first: the smaller radices go to the first passes.  For a zero-padded (LDE) input the first pass skips
// z = log2(blow-up) levels, and 2^23 = 2^7 2^8 2^8 with z = 3 leaves 4 levels = two radix-4 phases there and no radix-2
// phase anywhere, where 2^8 2^8 2^7 has one in the first and one in the last pass.
inline NttPlan ntt_make_plan(int log_n, bool small_first = false) {
    NttPlan pl;
    pl.log_n = log_n;
    int p = (log_n + ZKB_NTT_MAX_LOG_R - 1) / ZKB_NTT_MAX_LOG_R;
    if (p < 1) p = 1;
    pl.n_passes = p;
    int base = log_n / p, rem = log_n % p;
    for (int i = 0; i < ZKB_NTT_MAX_PASSES; i++) pl.lr[i] = 0;
    for (int i = 0; i < p; i++) pl.lr[i] = base + ((small_first ? i >= p - rem : i < rem) ? 1 : 0);
    // multi-pass tiles need R_1 >= C and M_i >= C (C = 8): true for every log_n with the default
    // ZKB_NTT_MAX_LOG_R = 8 (p >= 2 only when log_n >= 9 -> every lr >= 4).
    if (p >= 2 && (pl.lr[p - 1] < 3 || pl.lr[0] < 3)) pl.n_passes = -1;
    return pl;
}

// Zero-padded input (LDE): when the valid prefix is 2^v elements = 2^(v - log_row_stride) whole tile rows of the first
// pass, the first z = log_r - (v - log_row_stride) butterfly levels of that pass only replicate the live rows.
inline int ntt_zero_levels(uint64_t in_valid_elems, int log_row_stride, int log_r) {
    if (in_valid_elems == 0 || (in_valid_elems & (in_valid_elems - 1))) return 0;
    int v = 0;
    while ((1ull << v) < in_valid_elems) v++;
    int log_live = v - log_row_stride;
    if (log_live < 0 || log_live >= log_r) return 0;
    return log_r - log_live;
}

// Pointers to the tables one transform needs (device pointers in the runtime, host in the replay).
struct NttTables {
    const void *tw;                              // master in-tile twiddles for this direction
    int tw_in_smem;                              // stage every tile's twiddles in shared memory
    const void *inter[ZKB_NTT_MAX_PASSES];       // inter[i] = T_{i+1}, i < n_passes-1
    const void *load_tab;  uint64_t load_mask;   // coset pre-scale (or null)
    const void *store_tab; uint64_t store_mask;  // post-scale of the last pass (or null)
};

// Builds the per-pass parameter blocks.  `work` is a scratch buffer of batch * N elements, used
// when n_passes >= 2 (pass 1: in -> work, middle passes in place on work, last pass work -> out).
inline std::vector<NttPassParams> ntt_build_passes(const NttPlan &pl, const NttTables &tb, const void *in,
                                                   void *out, void *work, uint32_t batch,
                                                   uint64_t in_poly_stride, uint64_t out_poly_stride,
                                                   uint64_t in_valid_elems, const void *known_src = nullptr,
                                                   int known_log = 0, uint64_t known_poly_stride = 0) {
    std::vector<NttPassParams> v;
    const uint64_t N = 1ull << pl.log_n;
    const int p = pl.n_passes;
    // known outputs (see NttPassParams::known_log): multi-pass plans whose first radix covers the period, a whole
    // number of periods per column block of the last pass, and no post-scale on the last pass
    // (the kept k_1 values must fill whole column blocks of the last pass: R_1 / 2^known_log a multiple of 8)
    if (!(known_src && p >= 2 && known_log >= 3 && known_log + 3 <= pl.lr[0] && tb.store_tab == nullptr)) known_log = 0;
    for (int i = 0; i < p; i++) {
        NttPassParams q;
        q.batch = batch;
        q.log_r = pl.lr[i];
        q.log_n = pl.log_n;
        q.n_passes = p;
        for (int j = 0; j < ZKB_NTT_MAX_PASSES; j++) q.lr[j] = pl.lr[j];
        q.tw = tb.tw;
        q.tw_in_smem = tb.tw_in_smem;
        q.load_tab = nullptr; q.load_mask = 0;
        q.store_tab = nullptr; q.store_mask = 0;
        q.in_valid_elems = N;
        q.zero_levels = 0;
        q.known_log = known_log; q.known_k1_shift = 0;
        q.known_src = (const u128 *)known_src; q.known_poly_stride = known_poly_stride;
        q.log_m = 0; q.log_mprev = 0;
        if (p == 1) {
            q.mode = NTT_MODE_SINGLE;
            q.in = (const u128 *)in; q.out = (u128 *)out;
            q.in_poly_stride = in_poly_stride; q.out_poly_stride = out_poly_stride;
            q.tiles_per_poly = 0;
            q.in_valid_elems = in_valid_elems;
            q.zero_levels = ntt_zero_levels(in_valid_elems, 0, q.log_r);
            q.load_tab = tb.load_tab; q.load_mask = tb.load_mask;
            q.store_tab = tb.store_tab; q.store_mask = tb.store_mask;
        } else if (i < p - 1) {
            q.mode = NTT_MODE_STRIDED;
            q.log_m = pl.log_m(i + 1);
            q.log_mprev = pl.log_m(i);
            q.tiles_per_poly = N >> (q.log_r + 3);   // / (R * C), C = 8
            if (i >= 1 && known_log > 0) {           // the tiles of k_1 = 0 mod 2^known_log drop out of the grid
                q.known_k1_shift = (pl.log_n - q.log_mprev) - pl.lr[0];
                q.tiles_per_poly -= q.tiles_per_poly >> known_log;
            }
            if (i == 0) {
                q.in = (const u128 *)in; q.in_poly_stride = in_poly_stride;
                q.in_valid_elems = in_valid_elems;
                q.zero_levels = ntt_zero_levels(in_valid_elems, q.log_m, q.log_r);
                q.load_tab = tb.load_tab; q.load_mask = tb.load_mask;
            } else {
                q.in = (const u128 *)work; q.in_poly_stride = N;
            }
            q.out = (u128 *)work; q.out_poly_stride = N;
            q.store_tab = tb.inter[i];
            q.store_mask = (1ull << q.log_mprev) - 1;
        } else {
            q.mode = NTT_MODE_FINAL;
            q.log_m = pl.log_m(1);
            q.tiles_per_poly = N >> (q.log_r + 3);
            if (known_log > 0) q.tiles_per_poly -= q.tiles_per_poly >> known_log;   // no tile holds a known k_1
            q.in = (const u128 *)work; q.in_poly_stride = N;
            q.out = (u128 *)out; q.out_poly_stride = out_poly_stride;
            q.store_tab = tb.store_tab; q.store_mask = tb.store_mask;
        }
        v.push_back(q);
    }
    return v;
}

// out[poly][j 2^known_log] = known_src[poly][j]: element e of the flat range [0, batch * (N >> known_log)) (the device
// kernel in zkb_ntt.cu and the CPU replay both use this mapping)
ZKB_HD void ntt_known_scatter_index(uint64_t e, int log_n, int known_log, uint64_t known_poly_stride, uint64_t out_poly_stride,
                                    uint64_t *src, uint64_t *dst) {
    const uint64_t per_poly = 1ull << (log_n - known_log);
    const uint64_t poly = e / per_poly, j = e % per_poly;
    *src = poly * known_poly_stride + j;
    *dst = poly * out_poly_stride + (j << known_log);
}

inline uint64_t ntt_pass_tiles(const NttPassParams &q) {
    if (q.mode == NTT_MODE_SINGLE) return (q.batch + ZKB_NTT_C - 1) / ZKB_NTT_C;
    return q.tiles_per_poly * q.batch;
}

}  // namespace zkb
