// One pass of the multi-pass NTT: a CTA owns a tile of R = 2^log_r rows x ZKB_NTT_C columns held in
// shared memory, runs a radix-2^log_r sub-NTT down the rows (natural order in, bit-reversed out; radix-4 steps,
// 2 butterfly levels per shared-memory round trip), and writes the tile back multiplied by the
// inter-pass twiddles.  Columns are independent transforms, laid out so that every global access is
// a contiguous run (ZKB_NTT_C elements = 256 B along the column axis, or R elements when the row
// axis is the contiguous one) and every shared-memory access of a quarter-warp hits distinct banks.
//
// Device replacement for crypto3-math `basic_radix2_domain::fft / inverse_fft`
// (call sites: zk/snark/reductions/r1cs_to_qap.hpp:250-310,
//  zk/commitments/detail/polynomial/basic_fri.hpp:453) - same DFT, natural order in and out.
//
// The body is written as per-thread "phase" functions (load / step / store) that are
// __host__ __device__: the kernel calls them between __syncthreads(), and the CPU test harness
// (host_selftest.cpp) replays them thread by thread over a host buffer, which is how the index
// logic is validated in the GPU-less build container.
#pragma once
#include "zkb_field.cuh"

namespace zkb {

#define ZKB_NTT_C 8            // columns per tile
#ifndef ZKB_NTT_MAX_LOG_R
#define ZKB_NTT_MAX_LOG_R 8    // rows per tile <= 256 (overridable so the CPU replay can exercise 3-4 pass plans at small sizes)
#endif
#define ZKB_NTT_TW_LOG 10      // master in-tile twiddle table: w_{1024}^j, j < 512
#define ZKB_NTT_THREADS 256
#define ZKB_NTT_MAX_PASSES 4

struct alignas(16) u128 {
    uint32_t x, y, z, w;
};

enum { NTT_MODE_SINGLE = 0, NTT_MODE_STRIDED = 1, NTT_MODE_FINAL = 2 };

struct NttPassParams {
    const u128 *in;
    u128 *out;
    uint64_t in_poly_stride, out_poly_stride;  // elements between consecutive polynomials
    uint64_t tiles_per_poly;                   // SINGLE: tiles cover the batch, tiles_per_poly = 0
    uint32_t batch;
    int mode;
    int log_r;             // this pass' radix
    int log_n;             // full transform size
    int log_m;             // STRIDED: log2 M_i (row stride); FINAL: log2 M_1 (column stride)
    int log_mprev;         // STRIDED: log2 M_{i-1}
    int n_passes;
    int lr[ZKB_NTT_MAX_PASSES];   // log radix of every pass (for the FINAL digit reversal)
    uint64_t in_valid_elems;      // input indices (within the polynomial) >= this read as zero (zero-padded LDE)
    int zero_levels;              // z: only the first R >> z rows of every tile are non-zero (in_valid_elems is that
                                  // many whole rows).  A butterfly (a, b) -> (a + c b, a - c b) with b = 0 copies a, so
                                  // the first z levels only replicate the live rows: the load/premul phase writes the
                                  // replicas and the level loop starts z levels further down (no products for them).
    const void *load_tab;         // optional: in[i] *= load_tab[i & load_mask]   (Montgomery form)
    uint64_t load_mask;
    const void *store_tab;        // optional: out[o] *= store_tab[o & store_mask]
    uint64_t store_mask;
    const void *tw;               // master twiddles w_{2^ZKB_NTT_TW_LOG}^j (direction specific)
    int tw_in_smem;               // the tile's R/2 twiddles are staged in shared memory by the load phase
    // LDE: the outputs at multiples of 2^known_log are the input evaluations themselves (the coset of index 0 of the
    // larger domain is the smaller domain), out[poly][j 2^known_log] = known_src[poly][j].  Such outputs have
    // k_1 = 0 mod 2^known_log, so they are never computed: pass 1 does not store those rows, the middle passes skip
    // the tiles of those k_1, and the tiles of the last pass take their 8 columns from the kept k_1 only (one gap per
    // tile at most); ntt_known_scatter writes them from known_src.
    int known_log;                // 0 = off; otherwise 3 <= known_log <= lr[0]
    int known_k1_shift;           // middle pass: k_1 = q >> known_k1_shift
    const u128 *known_src;
    uint64_t known_poly_stride;
};

// shared-memory layout: two planes (low/high 16 bytes of every element), [row][C] 16-byte slots with the column index
// XOR-swizzled by the row (so that a quarter-warp hits eight distinct 16-byte bank groups whether its eight lanes walk
// along a row - butterflies, column-contiguous copies - or down a column - row-contiguous copies of the last pass),
// second plane offset by 4 extra slots so that the two halves of an element fall in complementary bank groups when a
// quarter-warp copies four elements x two halves.  64 KB + 64 B per 256-row tile (the padded [row][C+1] layout of round 1
// took 73 KB); after the planes: the tile's R/2 twiddles w_R^e as [e][2] slots (every quarter-warp reads one twiddle:
// a broadcast), staged once per CTA instead of fetched from global memory in every phase.
ZKB_HD constexpr uint32_t ntt_plane_slots(int log_r) { return (1u << log_r) * ZKB_NTT_C + 4; }
ZKB_HD constexpr uint32_t ntt_tw_slots(int log_r) { return log_r >= 1 ? (1u << log_r) : 2u; }      // (R/2) x 2
ZKB_HD constexpr uint32_t ntt_tw_base(int log_r) { return 2u * ntt_plane_slots(log_r); }
ZKB_HD constexpr uint32_t ntt_smem_bytes(int log_r) { return (2u * ntt_plane_slots(log_r) + ntt_tw_slots(log_r)) * 16u; }
ZKB_HD uint32_t ntt_slot(int log_r, int plane, uint32_t r, uint32_t c) {
    return plane * ntt_plane_slots(log_r) + r * ZKB_NTT_C + (c ^ (r & (ZKB_NTT_C - 1)));
}

ZKB_HD uint32_t brev(uint32_t v, int bits) {
#if defined(__CUDA_ARCH__)
    return bits ? (__brev(v) >> (32 - bits)) : 0;
#else
    uint32_t r = 0;
    for (int i = 0; i < bits; i++) r |= ((v >> i) & 1u) << (bits - 1 - i);
    return r;
#endif
}

struct NttTile {
    uint64_t in_base, out_base;                 // element offsets (including the polynomial offset)
    uint64_t in_row_stride, in_col_stride, out_row_stride, out_col_stride;
    uint64_t in_tab_base, out_tab_base;         // offsets within the polynomial (for the tables)
    uint32_t ncols;                             // valid columns
    uint32_t col_skip;                          // columns >= col_skip sit one k_1 further (the known k_1 between them is
                                                // not part of any tile); 0xffffffff: the 8 columns are consecutive
    uint64_t poly;
};

ZKB_HD NttTile ntt_tile(const NttPassParams &p, uint64_t tile) {
    NttTile t;
    const uint32_t C = ZKB_NTT_C;
    t.col_skip = 0xffffffffu;
    t.poly = 0;
    if (p.mode == NTT_MODE_SINGLE) {
        // columns = polynomials of the batch
        uint64_t b0 = tile * C;
        t.in_base = b0 * p.in_poly_stride;
        t.out_base = b0 * p.out_poly_stride;
        t.in_row_stride = 1; t.in_col_stride = p.in_poly_stride;
        t.out_row_stride = 1; t.out_col_stride = p.out_poly_stride;
        t.in_tab_base = 0; t.out_tab_base = 0;
        uint64_t left = p.batch - b0;
        t.ncols = left < C ? (uint32_t)left : C;
        return t;
    }
    // polynomial-minor order: the CTAs in flight at any time work on the same tile position of different polynomials, so
    // the slice of the N-entry inter-pass (or coset) table they share is read from DRAM once and then served by the L2
    // (polynomial-major order re-read the 268 MB table of a 2^23 transform for every polynomial)
    uint64_t poly = tile % p.batch;
    uint64_t tt = tile / p.batch;
    t.ncols = C;
    t.poly = poly;
    if (p.mode == NTT_MODE_STRIDED) {
        uint64_t blocks = (1ull << p.log_m) / C;       // column blocks per row group
        uint64_t q = tt / blocks, cb = tt % blocks;
        if (p.known_log > 0 && p.log_mprev < p.log_n) {
            // middle pass: the tiles of k_1 = 0 mod 2^known_log are not part of the grid
            uint64_t per_k1 = blocks << p.known_k1_shift;
            uint64_t k1p = tt / per_k1, rest = tt % per_k1;
            uint64_t k1 = k1p + k1p / ((1ull << p.known_log) - 1) + 1;
            q = (k1 << p.known_k1_shift) + rest / blocks;
            cb = rest % blocks;
        }
        uint64_t off = (q << p.log_mprev) + cb * C;
        t.in_tab_base = off; t.out_tab_base = off;
        t.in_base = poly * p.in_poly_stride + off;
        t.out_base = poly * p.out_poly_stride + off;
        t.in_row_stride = 1ull << p.log_m; t.in_col_stride = 1;
        t.out_row_stride = 1ull << p.log_m; t.out_col_stride = 1;
    } else {  // FINAL: rows contiguous on input, columns (k_1) contiguous on output
        uint64_t kblocks = (1ull << p.lr[0]) / C;
        uint64_t k1_0;                                 // k_1 of column 0
        if (p.known_log > 0) {
            // columns = 8 consecutive KEPT k_1 (those that are not 0 mod 2^known_log): kept index kappa -> k_1 = kappa +
            // kappa / per + 1 with per = 2^known_log - 1 kept values per period; 8 consecutive kappa cross one period
            // boundary at most (per >= 7)
            const uint64_t per = (1ull << p.known_log) - 1;
            kblocks = (((1ull << p.lr[0]) >> p.known_log) * per) / C;
            const uint64_t kappa0 = (tt % kblocks) * C;
            k1_0 = kappa0 + kappa0 / per + 1;
            const uint64_t c1 = per - kappa0 % per;
            if (c1 < C) t.col_skip = (uint32_t)c1;
        } else {
            k1_0 = (tt % kblocks) * C;
        }
        uint64_t q = tt / kblocks;                     // q = (k_2 .. k_{p-1}) mixed radix, k_2 major
        uint64_t in_off = (k1_0 << p.log_m) + (q << p.log_r);
        uint64_t rev = 0, qq = q;
        for (int i = p.n_passes - 2; i >= 1; i--) {
            uint64_t d = qq & ((1ull << p.lr[i]) - 1);
            qq >>= p.lr[i];
            rev = (rev << p.lr[i]) | d;
        }
        uint64_t out_off = k1_0 + (rev << p.lr[0]);
        t.in_tab_base = in_off; t.out_tab_base = out_off;
        t.in_base = poly * p.in_poly_stride + in_off;
        t.out_base = poly * p.out_poly_stride + out_off;
        t.in_row_stride = 1; t.in_col_stride = 1ull << p.log_m;
        t.out_row_stride = 1ull << (p.log_n - p.log_r); t.out_col_stride = 1;
    }
    return t;
}

template <class F>
ZKB_HD F ntt_ld_elem(const u128 *s, int log_r, uint32_t r, uint32_t c) {
    u128 lo = s[ntt_slot(log_r, 0, r, c)], hi = s[ntt_slot(log_r, 1, r, c)];
    F v;
    v.l[0] = lo.x; v.l[1] = lo.y; v.l[2] = lo.z; v.l[3] = lo.w;
    v.l[4] = hi.x; v.l[5] = hi.y; v.l[6] = hi.z; v.l[7] = hi.w;
    return v;
}
template <class F>
ZKB_HD void ntt_st_elem(u128 *s, int log_r, uint32_t r, uint32_t c, const F &v) {
    u128 lo = {v.l[0], v.l[1], v.l[2], v.l[3]}, hi = {v.l[4], v.l[5], v.l[6], v.l[7]};
    s[ntt_slot(log_r, 0, r, c)] = lo;
    s[ntt_slot(log_r, 1, r, c)] = hi;
}
template <class F>
ZKB_HD F ntt_ld_tab(const void *tab, uint64_t idx) {
    const u128 *t = (const u128 *)tab + 2 * idx;
    u128 lo = t[0], hi = t[1];
    F v;
    v.l[0] = lo.x; v.l[1] = lo.y; v.l[2] = lo.z; v.l[3] = lo.w;
    v.l[4] = hi.x; v.l[5] = hi.y; v.l[6] = hi.z; v.l[7] = hi.w;
    return v;
}

// ------------------------------------------------------------------------------------ load phase
// Copies the tile global -> shared (16-byte units, coalesced along whichever axis is contiguous).
template <class F>
ZKB_HD void ntt_phase_load(const NttPassParams &p, const NttTile &t, u128 *smem, uint32_t tid,
                                  uint32_t nthreads) {
    const uint32_t C = ZKB_NTT_C, R = 1u << p.log_r;
    const int log_l = p.log_r - p.zero_levels;
    const uint32_t L = 1u << log_l;                      // live rows (all of them unless zero_levels > 0)
    const uint32_t total = L * C * 2;  // 16-byte units
    const u128 zero = {0, 0, 0, 0};
    // with a coset pre-scale the premul phase writes the replicas (after scaling), otherwise the load does
    const uint32_t copies = p.load_tab != nullptr ? 1u : (1u << p.zero_levels);
    for (uint32_t u = tid; u < total; u += nthreads) {
        uint32_t half, r, c;
        if (t.in_col_stride == 1) {      // runs of C elements along the column axis
            half = u & 1; c = (u >> 1) % C; r = u / (2 * C);
        } else {                         // runs of L elements along the row axis (L = 2^log_l: shifts, not divisions)
            half = u & 1; r = (u >> 1) & (L - 1); c = u >> (log_l + 1);
        }
        const uint32_t ce = c + (c >= t.col_skip ? 1u : 0u);
        uint64_t widx = t.in_tab_base + r * t.in_row_stride + (p.mode == NTT_MODE_SINGLE ? 0 : ce * t.in_col_stride);
        const bool live = c < t.ncols && widx < p.in_valid_elems;
#if defined(__CUDA_ARCH__)
        // cp.async: global -> shared without a register round trip, so the 16 copies of a thread are all in flight at once
        // (a load followed by its own store serialised 16 DRAM round trips per thread: `long_scoreboard` on the STS was
        // 7 % of all stall samples, profiles/r3b_source.csv); src-size 0 writes zeros
        const u128 *src = live ? p.in + 2 * (t.in_base + r * t.in_row_stride + ce * t.in_col_stride) + half : p.in;
        const uint32_t nbytes = live ? 16u : 0u;
        for (uint32_t k = 0; k < copies; k++) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(smem + ntt_slot(p.log_r, half, r + k * L, c));
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
        }
#else
        u128 v = zero;
        if (live) v = p.in[2 * (t.in_base + r * t.in_row_stride + ce * t.in_col_stride) + half];
        for (uint32_t k = 0; k < copies; k++) smem[ntt_slot(p.log_r, half, r + k * L, c)] = v;
#endif
    }
#if defined(__CUDA_ARCH__)
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
#endif
    (void)zero;
}

// optional pre-multiplication (coset shift) - separate phase so that it works on whole elements
template <class F>
ZKB_HD void ntt_phase_premul(const NttPassParams &p, const NttTile &t, u128 *smem, uint32_t tid,
                                    uint32_t nthreads) {
    const uint32_t C = ZKB_NTT_C, R = 1u << p.log_r;
    const uint32_t L = R >> p.zero_levels, copies = 1u << p.zero_levels;
    for (uint32_t e = tid; e < L * C; e += nthreads) {
        uint32_t c = e % C, r = e / C;
        const uint32_t ce = c + (c >= t.col_skip ? 1u : 0u);
        uint64_t widx = t.in_tab_base + r * t.in_row_stride + (p.mode == NTT_MODE_SINGLE ? 0 : ce * t.in_col_stride);
        F v = ntt_ld_elem<F>(smem, p.log_r, r, c);
        if (c < t.ncols && widx < p.in_valid_elems) {
            uint64_t idx = widx & p.load_mask;
            v = v * ntt_ld_tab<F>(p.load_tab, idx);
        } else if (copies == 1) {
            continue;
        }
        for (uint32_t k = 0; k < copies; k++) ntt_st_elem<F>(smem, p.log_r, r + k * L, c, v);
    }
}

// ------------------------------------------------------------------------------------ butterflies
// stage the tile's twiddles w_R^e, e < R/2, into shared memory (one 16-byte unit per thread and step)
template <class F>
ZKB_HD void ntt_phase_stage_tw(const NttPassParams &p, u128 *smem, uint32_t tid, uint32_t nthreads) {
    const uint32_t units = ntt_tw_slots(p.log_r);
    const u128 *tab = (const u128 *)p.tw;
    u128 *dst = smem + ntt_tw_base(p.log_r);
    for (uint32_t u = tid; u < units; u += nthreads) {
        const uint32_t e = u >> 1;
        dst[u] = tab[2 * ((uint64_t)e << (ZKB_NTT_TW_LOG - p.log_r)) + (u & 1)];
    }
}
template <class F>
ZKB_HD F ntt_tw(const NttPassParams &p, const u128 *smem, uint32_t e, int log_size) {
    // w_{2^log_size}^e: from the staged copy (log_size = log_r always) or the master table
    if (p.tw_in_smem) {
        const u128 *t = smem + ntt_tw_base(p.log_r) + 2 * e;
        u128 lo = t[0], hi = t[1];
        F v;
        v.l[0] = lo.x; v.l[1] = lo.y; v.l[2] = lo.z; v.l[3] = lo.w;
        v.l[4] = hi.x; v.l[5] = hi.y; v.l[6] = hi.z; v.l[7] = hi.w;
        return v;
    }
    return ntt_ld_tab<F>(p.tw, (uint64_t)e << (ZKB_NTT_TW_LOG - log_size));
}

// The in-tile sub-transform splits the polynomial by halves (natural order in, bit-reversed order out)
// with Cooley-Tukey butterflies (a, b) -> (a + c b, a - c b): at the level with half-distance h = 2^log_h
// the rows form m = R / 2h blocks and block B evaluates on the coset w^{brev(B)} of the subgroup of order 2h,
// so its twiddle is c_B = w_R^{brev(B) h} - one constant per block, nothing per position.  Block 0 has c = 1
// (no product at all); compared with Gentleman-Sande butterflies (a + b, (a - b) w^j) this removes the unit
// twiddles of every level, not only of the last one.

// one level with half-distance h = 2^log_h
template <class F>
ZKB_HD void ntt_phase_radix2(const NttPassParams &p, u128 *smem, int log_h, uint32_t tid,
                                    uint32_t nthreads) {
    const uint32_t C = ZKB_NTT_C, R = 1u << p.log_r, h = 1u << log_h;
    const int lvl = p.log_r - 1 - log_h;               // number of blocks = 2^lvl
    for (uint32_t task = tid; task < (R / 2) * C; task += nthreads) {
        uint32_t c = task % C, g = task / C;
        uint32_t j = g & (h - 1), blk = g >> log_h;
        uint32_t i = (blk << (log_h + 1)) + j;
        F a = ntt_ld_elem<F>(smem, p.log_r, i, c), b = ntt_ld_elem<F>(smem, p.log_r, i + h, c);
        if (blk != 0) b = b * ntt_tw<F>(p, smem, brev(blk, lvl) << log_h, p.log_r);
        ntt_st_elem<F>(smem, p.log_r, i, c, a + b);
        ntt_st_elem<F>(smem, p.log_r, i + h, c, a - b);
    }
}

// two fused levels with half-distances h = 2^log_h and q = h/2: block B of the first level splits into
// blocks 2B, 2B+1 of the second; with beta = brev(B): c = w^{2 beta q}, c0 = w^{beta q}, c1 = c0 w_4
template <class F>
ZKB_HD void ntt_phase_radix4(const NttPassParams &p, u128 *smem, int log_h, uint32_t tid,
                                    uint32_t nthreads) {
    const uint32_t C = ZKB_NTT_C, R = 1u << p.log_r, h = 1u << log_h, q = h >> 1;
    const int log_q = log_h - 1;
    const int lvl = p.log_r - 1 - log_h;
    for (uint32_t task = tid; task < (R / 4) * C; task += nthreads) {
        uint32_t c = task % C, g = task / C;
        uint32_t j = g & (q - 1), blk = g >> log_q;
        uint32_t i = (blk << (log_h + 1)) + j;
        F x0 = ntt_ld_elem<F>(smem, p.log_r, i, c), x2 = ntt_ld_elem<F>(smem, p.log_r, i + h, c);
        F x1 = ntt_ld_elem<F>(smem, p.log_r, i + q, c), x3 = ntt_ld_elem<F>(smem, p.log_r, i + h + q, c);
        const uint32_t e0 = brev(blk, lvl) << log_q;    // beta q  (< R/4)
        if (blk != 0) {
            F w = ntt_tw<F>(p, smem, 2 * e0, p.log_r);
            x2 = x2 * w;
            x3 = x3 * w;
        }
        F y0 = x0 + x2, y2 = x0 - x2, y1 = x1 + x3, y3 = x1 - x3;
        if (blk != 0) y1 = y1 * ntt_tw<F>(p, smem, e0, p.log_r);
        y3 = y3 * ntt_tw<F>(p, smem, e0 + (R >> 2), p.log_r);
        ntt_st_elem<F>(smem, p.log_r, i, c, y0 + y1);
        ntt_st_elem<F>(smem, p.log_r, i + q, c, y0 - y1);
        ntt_st_elem<F>(smem, p.log_r, i + h, c, y2 + y3);
        ntt_st_elem<F>(smem, p.log_r, i + h + q, c, y2 - y3);
    }
}

// ------------------------------------------------------------------------------------ store phase
// Output row k lives in shared row bitrev(k) (DIF leaves the sub-transform bit-reversed).
template <class F>
ZKB_HD void ntt_phase_store(const NttPassParams &p, const NttTile &t, const u128 *smem, uint32_t tid,
                                   uint32_t nthreads) {
    const uint32_t C = ZKB_NTT_C, R = 1u << p.log_r;
    if (p.store_tab == nullptr) {
        const uint32_t total = R * C * 2;
        for (uint32_t u = tid; u < total; u += nthreads) {
            uint32_t half, k, c;
            if (t.out_col_stride == 1) {
                half = u & 1; c = (u >> 1) % C; k = u / (2 * C);
            } else {
                half = u & 1; k = (u >> 1) & (R - 1); c = u >> (p.log_r + 1);
            }
            if (c >= t.ncols) continue;
            const uint32_t ce = c + (c >= t.col_skip ? 1u : 0u);
            p.out[2 * (t.out_base + k * t.out_row_stride + ce * t.out_col_stride) + half] =
                smem[ntt_slot(p.log_r, half, brev(k, p.log_r), c)];
        }
        return;
    }
    // first pass of a transform with known outputs: rows k_1 = 0 mod 2^known_log are never read again
    const bool kskip = p.known_log > 0 && p.mode == NTT_MODE_STRIDED && p.log_mprev == p.log_n;
    const uint32_t kmask = kskip ? (1u << p.known_log) - 1 : 0u;
    for (uint32_t e = tid; e < R * C; e += nthreads) {
        uint32_t k, c;
        if (t.out_col_stride == 1) { c = e % C; k = e / C; } else { k = e & (R - 1); c = e >> p.log_r; }
        if (c >= t.ncols || (kskip && (k & kmask) == 0)) continue;
        const uint32_t ce = c + (c >= t.col_skip ? 1u : 0u);
        uint64_t off = k * t.out_row_stride + (p.mode == NTT_MODE_SINGLE ? 0 : ce * t.out_col_stride);
        F v = ntt_ld_elem<F>(smem, p.log_r, brev(k, p.log_r), c);
        v = v * ntt_ld_tab<F>(p.store_tab, (t.out_tab_base + off) & p.store_mask);
        u128 lo = {v.l[0], v.l[1], v.l[2], v.l[3]}, hi = {v.l[4], v.l[5], v.l[6], v.l[7]};
        u128 *dst = p.out + 2 * (t.out_base + k * t.out_row_stride + ce * t.out_col_stride);
        dst[0] = lo;
        dst[1] = hi;
    }
}

// Sequence of phases shared by the kernel and the host replay.  `sync` is __syncthreads() on the
// device and a no-op marker on the host (the host replays a phase for all threads before moving on).
#define ZKB_NTT_FOR_EACH_PHASE(PHASE_LOAD, PHASE_PREMUL, PHASE_R2, PHASE_R4, PHASE_STORE, log_r, zero_levels, has_load_tab) \
    {                                                                                                   \
        PHASE_LOAD;                                                                                     \
        if (has_load_tab) { PHASE_PREMUL; }                                                             \
        int _lh = (log_r)-1-(zero_levels);                                                              \
        if ((_lh + 1)&1) { PHASE_R2(_lh); _lh -= 1; }                                                   \
        for (; _lh >= 1; _lh -= 2) { PHASE_R4(_lh); }                                                   \
        PHASE_STORE;                                                                                    \
    }

#if defined(__CUDACC__)
template <class P>
__global__ void __launch_bounds__(ZKB_NTT_THREADS, 3) ntt_pass_kernel(const NttPassParams p) {
    typedef Fp<P> F;
    extern __shared__ uint4 ntt_smem_raw[];
    u128 *smem = reinterpret_cast<u128 *>(ntt_smem_raw);
    const NttTile t = ntt_tile(p, blockIdx.x);
    const uint32_t tid = threadIdx.x, nt = blockDim.x;
#define ZKB_L if (p.tw_in_smem) ntt_phase_stage_tw<F>(p, smem, tid, nt); ntt_phase_load<F>(p, t, smem, tid, nt); __syncthreads()
#define ZKB_PM ntt_phase_premul<F>(p, t, smem, tid, nt); __syncthreads()
#define ZKB_R2(lh) ntt_phase_radix2<F>(p, smem, lh, tid, nt); __syncthreads()
#define ZKB_R4(lh) ntt_phase_radix4<F>(p, smem, lh, tid, nt); __syncthreads()
#define ZKB_S ntt_phase_store<F>(p, t, smem, tid, nt)
    ZKB_NTT_FOR_EACH_PHASE(ZKB_L, ZKB_PM, ZKB_R2, ZKB_R4, ZKB_S, p.log_r, p.zero_levels, p.load_tab != nullptr)
#undef ZKB_L
#undef ZKB_PM
#undef ZKB_R2
#undef ZKB_R4
#undef ZKB_S
}
#endif

}  // namespace zkb
