// Pippenger multi-scalar multiplication on G1 (BLS12-381, BN254, Pallas) for sm_100a.
//
// Device replacement for `algebra::multiexp<Method>` / `multiexp_with_mixed_addition<Method>`
// (call sites: zk/commitments/polynomial/kzg.hpp:146, zk/snark/systems/ppzksnark/r1cs_gg_ppzksnark/
// prover.hpp:108-139, zk/commitments/polynomial/knowledge_commitment_multiexp.hpp:107).  The result is
// the same group element whatever the method; parity is checked in affine form.
//
// Pipeline (W = ceil((bits+1)/c) signed c-bit windows, 2^(c-1) buckets per window):
//   1 digits      scalar -> W signed digits; key = (bucket << 1 | sign); histogram of bucket sizes
//   2 scan        bucket offsets; split buckets longer than MSM_TASK_CAP into tasks (0/1-heavy
//                 Groth16 assignments put most points into one bucket - the reference pre-filters
//                 them on the CPU, knowledge_commitment_multiexp.hpp:88-101)
//   3 scatter     counting sort of (point index, sign) by bucket
//   4 order       tasks are ordered by length (counting sort over <= MSM_TASK_CAP lengths) so that the 32
//                 tasks of a warp run the same number of additions (bucket sizes are Poisson distributed;
//                 unsorted, a warp waits for its longest bucket: ~30 % of the lanes idle at 32 points/bucket)
//   5 accumulate  one thread per task: XYZZ += affine point (mixed add, 8M+2S)
//   6 reduce      per window S_w = sum_k (k+1) B_k by a radix-MSM_RED_L tree: every node turns L children
//                 (A_j = plain sum, U_j = weighted sum relative to the child's start) into
//                 A = sum A_j, U = sum U_j + span * sum j A_j  - short running sums at every level,
//                 65536 threads at the leaves, no serial pass over the 2^(c-1) buckets of a window
//   7 combine     sum_w 2^(c w) S_w on the host (c W doublings; zkb_msm_host.cpp)
#include <stdio.h>
#include <string.h>
#ifndef ZKB_MSM_MUL_INLINE
#define ZKB_MUL_OUTLINE 1
#endif
#include "zkb_curve.cuh"
#include "zkb_internal.h"

using namespace zkb;

namespace zkb {
int msm_window_combine(int curve, int c, int W, const uint32_t *sums, uint32_t *out);
}

struct zkb_msm_bases {
    int device;      // bases are plain device memory: any context of the same device may use them (one context per
                     // concurrent call - a context's scratch serves one call at a time)
    int curve;
    uint64_t n;
    void *d_points;  // Affine<F>, Montgomery form
    // optional window table (zkb_msm_bases_precompute): table[w * n + i] = 2^(table_c * w) * P_i, affine,
    // Montgomery form, w < table_w; window 0 is d_points itself when table == nullptr
    void *d_table = nullptr;
    int table_c = 0, table_w = 0;
};

// coordinate fields of the supported groups
typedef Fp<params::Bls12381Fq> FqBls;
typedef Fp<params::Bn254Fq> FqBn;
typedef Fp<params::PallasFp> FpPallas;
typedef Fp2<FqBls> Fq2Bls;
typedef Fp2<FqBn> Fq2Bn;
#define ZKB_DISPATCH_CURVE(curve, ...)                                 \
    switch (curve) {                                                     \
        case ZKB_CURVE_BLS12_381_G1: { typedef FqBls CF; enum { SB = params::Bls12381Fr::BITS }; __VA_ARGS__; } break;  \
        case ZKB_CURVE_BN254_G1: { typedef FqBn CF; enum { SB = params::Bn254Fr::BITS }; __VA_ARGS__; } break;          \
        case ZKB_CURVE_PALLAS: { typedef FpPallas CF; enum { SB = params::PallasFq::BITS }; __VA_ARGS__; } break;       \
        case ZKB_CURVE_BLS12_381_G2: { typedef Fq2Bls CF; enum { SB = params::Bls12381Fr::BITS }; __VA_ARGS__; } break; \
        case ZKB_CURVE_BN254_G2: { typedef Fq2Bn CF; enum { SB = params::Bn254Fr::BITS }; __VA_ARGS__; } break;         \
        default: break;                                                  \
    }

#define MSM_TASK_CAP 256u
#define MSM_HEAVY_MIN_TASKS 8u
#define MSM_SENTINEL 0xffffffffu
#define MSM_LEVEL_BITS 5

// ------------------------------------------------------------------------------------ small kernels
template <class F>
__global__ void __launch_bounds__(256) points_to_mont_kernel(uint64_t n, Affine<F> *pts) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) pts[i] = pts[i].to_mont();
}

// scalars: 8 canonical limbs each.  keys[w * n + i]
// win_stride = buckets per window (2^(c-1)), or 0 when all windows share one bucket set (window table)
__global__ void __launch_bounds__(256) msm_digits_kernel(uint32_t n, const uint32_t *__restrict__ scalars, int c, int W,
                                                         uint32_t win_stride, uint32_t *__restrict__ keys,
                                                         uint32_t *__restrict__ counts) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4 *sp = reinterpret_cast<const uint4 *>(scalars) + 2 * (uint64_t)i;
    uint4 lo = sp[0], hi = sp[1];
    uint32_t l[9] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w, 0};
    const uint32_t half = 1u << (c - 1), mask = (1u << c) - 1;
    uint32_t carry = 0;
    for (int w = 0; w < W; w++) {
        uint32_t bit = w * c, limb = bit >> 5, sh = bit & 31;
        uint64_t two = limb < 8 ? ((uint64_t)l[limb] | ((uint64_t)l[limb + 1] << 32)) : 0;
        uint32_t raw = ((uint32_t)(two >> sh) & mask) + carry;
        uint32_t key = MSM_SENTINEL;
        carry = 0;
        if (raw != 0) {
            uint32_t mag = raw, neg = 0;
            if (raw > half && w < W - 1) {  // recode into [-2^(c-1), 2^(c-1)]
                mag = (1u << c) - raw;
                neg = 1;
                carry = 1;
            }
            if (mag != 0) {
                uint32_t bucket = (uint32_t)w * win_stride + (mag - 1);
                key = (bucket << 1) | neg;
                atomicAdd(counts + bucket, 1u);
            }
        }
        keys[(uint64_t)w * n + i] = key;
    }
}

// tasks per bucket from the bucket sizes; buckets split into more than MSM_HEAVY_MIN_TASKS tasks are listed in `heavy`
__global__ void __launch_bounds__(256) msm_ntasks_kernel(uint32_t nb, const uint32_t *__restrict__ counts,
                                                         uint32_t *__restrict__ ntasks, uint32_t *__restrict__ n_heavy,
                                                         uint32_t *__restrict__ heavy, uint32_t heavy_cap) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    uint32_t nt = (counts[b] + MSM_TASK_CAP - 1) / MSM_TASK_CAP;
    ntasks[b] = nt;
    if (nt > MSM_HEAVY_MIN_TASKS) {   // a few tasks are added serially by the level-0 reduce
        uint32_t k = atomicAdd(n_heavy, 1u);
        if (k < heavy_cap) heavy[k] = b;
    }
}

// ---- exclusive scan over uint32 (three launches, 1024 items per block) --------------------------
__global__ void __launch_bounds__(256) scan_block_kernel(uint32_t n, const uint32_t *__restrict__ in,
                                                         uint32_t *__restrict__ out, uint32_t *__restrict__ block_sums) {
    __shared__ uint32_t warp_sums[8];
    uint32_t base = blockIdx.x * 1024 + threadIdx.x * 4;
    uint32_t v[4], s = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        v[k] = base + k < n ? in[base + k] : 0;
        s += v[k];
    }
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t incl = s;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
    }
    if (lane == 31) warp_sums[wid] = incl;
    __syncthreads();
    uint32_t woff = 0;
    for (uint32_t k = 0; k < wid; k++) woff += warp_sums[k];
    uint32_t excl = woff + incl - s;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        if (base + k < n) out[base + k] = excl;
        excl += v[k];
    }
    if (threadIdx.x == 255) block_sums[blockIdx.x] = woff + incl;
}
__global__ void __launch_bounds__(1024) scan_sums_kernel(uint32_t nblocks, uint32_t *block_sums, uint32_t *total) {
    // single block: serial over chunks of 1024 block sums
    __shared__ uint32_t sh[1024];
    __shared__ uint32_t running;
    if (threadIdx.x == 0) running = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nblocks; base += 1024) {
        uint32_t i = base + threadIdx.x;
        uint32_t v = i < nblocks ? block_sums[i] : 0;
        sh[threadIdx.x] = v;
        __syncthreads();
        for (uint32_t d = 1; d < 1024; d <<= 1) {
            uint32_t t = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
            __syncthreads();
            sh[threadIdx.x] += t;
            __syncthreads();
        }
        uint32_t r = running;
        if (i < nblocks) block_sums[i] = r + sh[threadIdx.x] - v;
        __syncthreads();
        if (threadIdx.x == 1023) running = r + sh[1023];
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = running;
}
__global__ void __launch_bounds__(256) scan_add_kernel(uint32_t n, uint32_t *__restrict__ out,
                                                       const uint32_t *__restrict__ block_sums) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += block_sums[i >> 10];
}

static int exclusive_scan(zkb_ctx *ctx, uint32_t n, const uint32_t *in, uint32_t *out, uint32_t *block_sums,
                          uint32_t *total, cudaStream_t st) {
    uint32_t nblocks = (n + 1023) / 1024;
    scan_block_kernel<<<nblocks, 256, 0, st>>>(n, in, out, block_sums);
    scan_sums_kernel<<<1, 1024, 0, st>>>(nblocks, block_sums, total);
    scan_add_kernel<<<(n + 255) / 256, 256, 0, st>>>(n, out, block_sums);
    ctx->launches += 3;
    ZKB_CUDA_OK(ctx, cudaGetLastError());
    return ZKB_OK;
}

struct MsmTask {
    uint32_t start;   // offset into the sorted index array
    uint32_t len;
};

// one thread per bucket: emit its tasks
__global__ void __launch_bounds__(256) msm_fill_tasks_kernel(uint32_t nb, const uint32_t *__restrict__ counts,
                                                             const uint32_t *__restrict__ offsets,
                                                             const uint32_t *__restrict__ task_offsets, MsmTask *tasks) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    uint32_t cnt = counts[b], off = offsets[b], t = task_offsets[b];
    for (uint32_t s = 0; s < cnt; s += MSM_TASK_CAP, t++) {
        MsmTask k;
        k.start = off + s;
        k.len = cnt - s < MSM_TASK_CAP ? cnt - s : MSM_TASK_CAP;
        tasks[t] = k;
    }
}

// pt_stride = 0: every window reads the same n points; window table: window w reads points [w * pt_stride, ..)
__global__ void __launch_bounds__(256) msm_scatter_kernel(uint64_t total, uint32_t n, uint32_t pt_stride,
                                                          const uint32_t *__restrict__ keys,
                                                          const uint32_t *__restrict__ offsets, uint32_t *__restrict__ cursor,
                                                          uint32_t *__restrict__ sorted) {
    uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    uint32_t key = keys[idx];
    if (key == MSM_SENTINEL) return;
    uint32_t bucket = key >> 1;
    uint32_t i = (uint32_t)(idx % n) + (uint32_t)(idx / n) * pt_stride;
    uint32_t pos = atomicAdd(cursor + bucket, 1u);
    sorted[offsets[bucket] + pos] = (i << 1) | (key & 1);
}

// ------------------------------------------------------------------------------------ accumulate
template <class F>
__device__ __forceinline__ Affine<F> load_affine(const Affine<F> *p) {
    // 2*N limbs, 16-byte aligned (N = 8 or 12): vector loads
    Affine<F> r;
    const uint4 *s = reinterpret_cast<const uint4 *>(p);
    uint32_t *d = reinterpret_cast<uint32_t *>(&r);
#pragma unroll
    for (int k = 0; k < (2 * F::N) / 4; k++) {
        uint4 v = __ldg(s + k);
        d[4 * k] = v.x; d[4 * k + 1] = v.y; d[4 * k + 2] = v.z; d[4 * k + 3] = v.w;
    }
    return r;
}

template <class F>
__global__ void __launch_bounds__(128) msm_accumulate_kernel(const uint32_t *__restrict__ n_tasks, const MsmTask *__restrict__ tasks,
                                                             const uint32_t *__restrict__ perm, const uint32_t *__restrict__ sorted,
                                                             const Affine<F> *__restrict__ points,
                                                             XYZZ<F> *__restrict__ out) {
    uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= *n_tasks) return;
    const uint32_t t = perm[g];
    MsmTask task = tasks[t];
    XYZZ<F> acc = XYZZ<F>::infinity();
    for (uint32_t k = 0; k < task.len; k++) {
        uint32_t e = sorted[task.start + k];
        Affine<F> p = load_affine<F>(points + (e >> 1));
        if (e & 1) p.y = p.y.neg();
        acc.add_mixed(p);
    }
    out[t] = acc;
}

// ------------------------------------------------------------------------------------ task order
#define MSM_RED_LOG_L 3
#define MSM_RED_L (1u << MSM_RED_LOG_L)

// bins[MSM_TASK_CAP - len]++ (descending length order)
__global__ void __launch_bounds__(256) msm_task_hist_kernel(const uint32_t *__restrict__ n_tasks, const MsmTask *__restrict__ tasks,
                                                            uint32_t *__restrict__ bins) {
    __shared__ uint32_t sh[MSM_TASK_CAP];
    for (uint32_t k = threadIdx.x; k < MSM_TASK_CAP; k += blockDim.x) sh[k] = 0;
    __syncthreads();
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t < *n_tasks) atomicAdd(&sh[MSM_TASK_CAP - tasks[t].len], 1u);
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < MSM_TASK_CAP; k += blockDim.x)
        if (sh[k]) atomicAdd(bins + k, sh[k]);
}
// single block: exclusive scan of the MSM_TASK_CAP bins into bin_off; clears the cursors
__global__ void __launch_bounds__(MSM_TASK_CAP) msm_task_scan_kernel(const uint32_t *__restrict__ bins, uint32_t *__restrict__ bin_off,
                                                                     uint32_t *__restrict__ cursor) {
    __shared__ uint32_t sh[MSM_TASK_CAP];
    uint32_t v = bins[threadIdx.x];
    sh[threadIdx.x] = v;
    __syncthreads();
    for (uint32_t d = 1; d < MSM_TASK_CAP; d <<= 1) {
        uint32_t t = threadIdx.x >= d ? sh[threadIdx.x - d] : 0;
        __syncthreads();
        sh[threadIdx.x] += t;
        __syncthreads();
    }
    bin_off[threadIdx.x] = sh[threadIdx.x] - v;
    cursor[threadIdx.x] = 0;
}
// perm[position in length order] = task id (one global atomic per block and non-empty bin)
__global__ void __launch_bounds__(256) msm_task_permute_kernel(const uint32_t *__restrict__ n_tasks, const MsmTask *__restrict__ tasks,
                                                               const uint32_t *__restrict__ bin_off, uint32_t *__restrict__ cursor,
                                                               uint32_t *__restrict__ perm) {
    __shared__ uint32_t cnt[MSM_TASK_CAP], base[MSM_TASK_CAP];
    for (uint32_t k = threadIdx.x; k < MSM_TASK_CAP; k += blockDim.x) cnt[k] = 0;
    __syncthreads();
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t bin = 0, rank = 0;
    bool live = t < *n_tasks;
    if (live) {
        bin = MSM_TASK_CAP - tasks[t].len;
        rank = atomicAdd(&cnt[bin], 1u);
    }
    __syncthreads();
    for (uint32_t k = threadIdx.x; k < MSM_TASK_CAP; k += blockDim.x)
        if (cnt[k]) base[k] = atomicAdd(cursor + k, cnt[k]);
    __syncthreads();
    if (live) perm[bin_off[bin] + base[bin] + rank] = t;
}

// ------------------------------------------------------------------------------------ heavy buckets
// A bucket with more than MSM_TASK_CAP points was accumulated as several tasks (the top window of a
// 255-bit scalar has only a few significant bits, so its handful of buckets hold N/8 points each; 0/1-heavy
// Groth16 assignments put almost everything into bucket 0 of window 0 - the reference filters those on the
// CPU, knowledge_commitment_multiexp.hpp:88-101).  One block per such bucket adds its task results pairwise
// in place (a tree over global memory), leaves the total in the bucket's first task and sets ntasks to 1.
template <class F>
__global__ void __launch_bounds__(256) msm_heavy_tree_kernel(const uint32_t *__restrict__ n_heavy, const uint32_t *__restrict__ heavy,
                                                             uint32_t heavy_cap, const uint32_t *__restrict__ task_offsets,
                                                             uint32_t *__restrict__ ntasks, XYZZ<F> *__restrict__ tout) {
    uint32_t nh = *n_heavy;
    if (nh > heavy_cap) nh = heavy_cap;
    for (uint32_t h = blockIdx.x; h < nh; h += gridDim.x) {
        const uint32_t b = heavy[h], t0 = task_offsets[b], nt = ntasks[b];
        for (uint32_t s = 1; s < nt; s <<= 1) {
            for (uint32_t i = threadIdx.x * 2 * s; i + s < nt; i += blockDim.x * 2 * s) {
                XYZZ<F> x = tout[t0 + i];
                x.add(tout[t0 + i + s]);
                tout[t0 + i] = x;
            }
            __syncthreads();
        }
        __syncthreads();
        if (threadIdx.x == 0) ntasks[b] = 1;
    }
}

// ------------------------------------------------------------------------------------ reduce
// One tree level.  Items of a window are (A, U) pairs (level 0: A = bucket value = sum of the bucket's task
// results, U = 0).  Node (w, s) folds items s*L .. s*L+L-1:
//     A' = sum_j A_j                                   ("run": running sum from the top item down)
//     U' = sum_j U_j + 2^span_bits * sum_j j A_j       ("acc" += run after every step but the last; "usum")
// so that at the root S_w = sum_k (k+1) B_k = U + A (written to outA when `root`).
// The three running sums of a node are three dependent chains of XYZZ additions, and the upper levels have
// far fewer nodes than the GPU has lanes, so a node is spread over LPN adjacent lanes of a warp (roles run /
// usum / acc; acc trails run by one step through shared memory): the chain per level is L + 1 additions
// instead of 3 L.  Level 0 has no U: LPN = 2.
template <class F, int LPN>
__global__ void __launch_bounds__(128) msm_reduce_level_kernel(int W, uint32_t n_in, int span_bits, int root,
                                                               const uint32_t *__restrict__ task_offsets,
                                                               const uint32_t *__restrict__ ntasks,
                                                               const XYZZ<F> *__restrict__ inA,
                                                               const XYZZ<F> *__restrict__ inU,
                                                               XYZZ<F> *__restrict__ outA, XYZZ<F> *__restrict__ outU) {
    typedef XYZZ<F> Pt;
    constexpr bool LEVEL0 = LPN == 2;
    constexpr int ROLE_RUN = 0, ROLE_ACC = 1, ROLE_USUM = 2;          // LPN == 4: lane 3 idles
    constexpr int NODES = 128 / LPN;
    __shared__ Pt sh_run[2][NODES];
    __shared__ Pt sh_usum[LEVEL0 ? 1 : NODES];
    const uint32_t n_out = (n_in + MSM_RED_L - 1) >> MSM_RED_LOG_L;
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t node = gtid / LPN, ln = threadIdx.x / LPN;
    const int role = gtid % LPN;
    const bool live = node < (uint32_t)W * n_out;
    const uint32_t w = live ? node / n_out : 0, s = live ? node % n_out : 0;
    const uint32_t j0 = s << MSM_RED_LOG_L;
    uint32_t j1 = j0 + MSM_RED_L;
    if (j1 > n_in) j1 = n_in;
    Pt x = Pt::infinity();
    for (uint32_t step = 0; step <= MSM_RED_L; step++) {
        // step k handles item j = j1-1-k (run, usum) and folds the run value of step k-1 into acc
        const bool has_item = live && step < j1 - j0;
        const uint32_t j = j1 - 1 - step;
        const uint64_t idx = (uint64_t)w * n_in + j;
        // every role funnels into ONE call site of add() (a warp runs its roles in lockstep)
        const Pt *src = nullptr;
        uint32_t cnt = 0;
        if (role == ROLE_RUN && has_item) {
            if (LEVEL0) { src = inA + task_offsets[idx]; cnt = ntasks[idx]; }
            else { src = inA + idx; cnt = 1; }
        } else if (role == ROLE_USUM && has_item) {
            src = inU + idx; cnt = 1;
        } else if (role == ROLE_ACC && live && step >= 1 && step < j1 - j0) {
            src = &sh_run[(step - 1) & 1][ln]; cnt = 1;   // run after item j1-step, which is > j0
        }
        for (uint32_t t = 0; t < cnt; t++) x.add(src[t]);
        if (role == ROLE_RUN && has_item) sh_run[step & 1][ln] = x;
        __syncwarp();
    }
    if (!LEVEL0 && role == ROLE_USUM && live) sh_usum[ln] = x;
    __syncwarp();
    if (!live) return;
    if (role == ROLE_RUN) {
        if (!root) outA[node] = x;
    } else if (role == ROLE_ACC) {
        for (int i = 0; i < span_bits; i++) x = x.dbl();
        if (!LEVEL0) x.add(sh_usum[ln]);
        if (root) {   // S_w = U + A
            x.add(sh_run[(j1 - j0 - 1) & 1][ln]);
            outA[node] = x;
        } else {
            outU[node] = x;
        }
    }
}

// ---- upper levels: one XYZZ operation spread over a quad of lanes ---------------------------------
// The upper tree levels have few nodes and are pure dependent chains: what counts is the latency of one
// XYZZ addition (14 sequential field products for a lone lane, ~16 us).  Here the four lanes of a quad hold
// the same operands, each computes one of up to four independent products of a stage, and the products
// are exchanged with warp shuffles: an addition is 4 product latencies deep, a doubling 3.
template <class P>
__device__ __forceinline__ Fp<P> quad_from(const Fp<P> &x, int src) {
    Fp<P> r;
#pragma unroll
    for (int i = 0; i < P::N; i++) r.l[i] = __shfl_sync(0xffffffffu, x.l[i], src, 4);
    return r;
}
template <class F>
__device__ __forceinline__ F quad_pick(int sub, const F &a0, const F &a1, const F &a2, const F &a3) {
    F r;
#pragma unroll
    for (int i = 0; i < F::N; i++) r.l[i] = sub == 0 ? a0.l[i] : sub == 1 ? a1.l[i] : sub == 2 ? a2.l[i] : a3.l[i];
    return r;
}
// p + q on every lane of the quad (all 32 lanes of the warp must call it; `sub` = lane & 3)
template <class F>
__device__ __forceinline__ XYZZ<F> quad_add(const XYZZ<F> &p, const XYZZ<F> &q, int sub) {
    const bool p_inf = p.is_infinity(), q_inf = q.is_infinity();
    F t = quad_pick(sub, p.X, q.X, p.Y, q.Y) * quad_pick(sub, q.ZZ, p.ZZ, q.ZZZ, p.ZZZ);
    const F U1 = quad_from(t, 0), U2 = quad_from(t, 1), S1 = quad_from(t, 2), S2 = quad_from(t, 3);
    const F Pd = U2 - U1, R = S2 - S1;
    t = quad_pick(sub, Pd, R, p.ZZ, p.ZZZ) * quad_pick(sub, Pd, R, q.ZZ, q.ZZZ);
    const F PP = quad_from(t, 0), RR = quad_from(t, 1), ZZa = quad_from(t, 2), ZZZa = quad_from(t, 3);
    t = quad_pick(sub, Pd, U1, ZZa, ZZa) * PP;
    const F PPP = quad_from(t, 0), Q = quad_from(t, 1);
    XYZZ<F> r;
    r.ZZ = quad_from(t, 2);
    r.X = RR - PPP - Q.dbl();
    t = quad_pick(sub, R, S1, ZZZa, ZZZa) * quad_pick(sub, Q - r.X, PPP, PPP, PPP);
    r.Y = quad_from(t, 0) - quad_from(t, 1);
    r.ZZZ = quad_from(t, 2);
    // exceptional inputs are uniform over the quad (same operands on its four lanes)
    if (q_inf) return p;
    if (p_inf) return q;
    if (Pd.is_zero()) return R.is_zero() ? p.dbl() : XYZZ<F>::infinity();
    return r;
}
// 2 p on every lane of the quad (dbl-2008-s-1)
template <class F>
__device__ __forceinline__ XYZZ<F> quad_dbl(const XYZZ<F> &p, int sub) {
    const F U = p.Y.dbl();
    F t = quad_pick(sub, U, p.X, U, U) * quad_pick(sub, U, p.X, U, U);
    const F V = quad_from(t, 0), XX = quad_from(t, 1);
    const F M = XX.dbl() + XX;
    t = quad_pick(sub, U, p.X, M, V) * quad_pick(sub, V, V, M, p.ZZ);
    const F W = quad_from(t, 0), S = quad_from(t, 1), MM = quad_from(t, 2);
    XYZZ<F> r;
    r.ZZ = quad_from(t, 3);
    r.X = MM - S.dbl();
    t = quad_pick(sub, M, W, W, W) * quad_pick(sub, S - r.X, p.Y, p.ZZZ, p.ZZZ);
    r.Y = quad_from(t, 0) - quad_from(t, 1);
    r.ZZZ = quad_from(t, 2);
    if (p.is_infinity() || p.Y.is_zero()) return XYZZ<F>::infinity();
    return r;
}

// Same node arithmetic as msm_reduce_level_kernel<F, 4> (roles run / acc / usum) with every role's accumulator
// replicated over a quad: 16 lanes per node, 8 nodes per 128-thread block.  Upper levels only (inA/inU are
// (A, U) arrays of the previous level).
template <class F>
__global__ void __launch_bounds__(128) msm_reduce_level_quad_kernel(int W, uint32_t n_in, int span_bits, int root,
                                                                    const XYZZ<F> *__restrict__ inA,
                                                                    const XYZZ<F> *__restrict__ inU,
                                                                    XYZZ<F> *__restrict__ outA, XYZZ<F> *__restrict__ outU) {
    typedef XYZZ<F> Pt;
    constexpr int ROLE_RUN = 0, ROLE_ACC = 1, ROLE_USUM = 2;   // role 3 idles
    constexpr int NODES = 128 / 16;
    __shared__ Pt sh_run[2][NODES];
    __shared__ Pt sh_usum[NODES];
    const uint32_t n_out = (n_in + MSM_RED_L - 1) >> MSM_RED_LOG_L;
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t node = gtid >> 4, ln = threadIdx.x >> 4;
    const int role = (gtid >> 2) & 3, sub = gtid & 3;
    const bool live = node < (uint32_t)W * n_out;
    const uint32_t w = live ? node / n_out : 0, s = live ? node % n_out : 0;
    const uint32_t j0 = s << MSM_RED_LOG_L;
    uint32_t j1 = j0 + MSM_RED_L;
    if (j1 > n_in) j1 = n_in;
    const Pt inf = Pt::infinity();
    Pt x = inf;
    for (uint32_t step = 0; step <= MSM_RED_L; step++) {
        const bool has_item = live && step < j1 - j0;
        const uint64_t idx = (uint64_t)w * n_in + (j1 - 1 - step);
        const Pt *src = nullptr;
        if (role == ROLE_RUN && has_item) src = inA + idx;
        else if (role == ROLE_USUM && has_item) src = inU + idx;
        else if (role == ROLE_ACC && live && step >= 1 && step < j1 - j0) src = &sh_run[(step - 1) & 1][ln];
        // every lane of the warp runs the quad addition (shuffles); lanes without work add infinity
        x = quad_add(x, src ? *src : inf, sub);
        if (role == ROLE_RUN && has_item && sub == 0) sh_run[step & 1][ln] = x;
        __syncwarp();
    }
    if (role == ROLE_USUM && live && sub == 0) sh_usum[ln] = x;
    __syncwarp();
    // acc: * 2^span_bits, + usum (+ run at the root); the other roles run along on infinity
    Pt y = role == ROLE_ACC ? x : inf;
    for (int i = 0; i < span_bits; i++) y = quad_dbl(y, sub);
    y = quad_add(y, role == ROLE_ACC && live ? sh_usum[ln] : inf, sub);
    if (root) y = quad_add(y, role == ROLE_ACC && live ? sh_run[(j1 - j0 - 1) & 1][ln] : inf, sub);
    if (!live || sub != 0) return;
    if (role == ROLE_RUN && !root) outA[node] = x;
    if (role == ROLE_ACC) (root ? outA : outU)[node] = y;
}

// ------------------------------------------------------------------------------------ host driver
static int msm_pick_c(uint64_t n) {
    int lg = 0;
    while ((1ull << lg) < n) lg++;
    int c = lg - 4;
    if (c < 2) c = 2;
    if (c > 20) c = 20;
    return c;
}

template <class F, int SCALAR_BITS>
static int msm_run_t(zkb_ctx *ctx, const zkb_msm_bases *bases, uint64_t offset, uint64_t n, const void *d_scalars,
                     uint32_t *partial_host, cudaStream_t st) {
    typedef XYZZ<F> Pt;
    // with a window table every digit window adds 2^(c w) P_i straight into ONE bucket set
    const bool tabled = bases->d_table != nullptr;
    const int c = tabled ? bases->table_c : msm_pick_c(n);
    const int W = (SCALAR_BITS + 1 + c - 1) / c;        // digit windows
    const int WB = tabled ? 1 : W;                      // bucket sets
    const uint32_t M = 1u << (c - 1);
    const uint32_t nb = (uint32_t)WB * M;
    if (tabled && (W > bases->table_w || (uint64_t)W * bases->n >= (1ull << 31)))
        return ctx_fail(ctx, ZKB_ERR_UNSUPPORTED, "MSM: window table does not cover the scalar width");
    const uint64_t total_keys = (uint64_t)W * n;
    if (n >= (1ull << 31) || total_keys >= (1ull << 32))
        return ctx_fail(ctx, ZKB_ERR_UNSUPPORTED, "MSM: W*n must be < 2^32 (split the range across calls/GPUs)");
    const uint64_t max_tasks = (uint64_t)nb + total_keys / MSM_TASK_CAP + 1;

    // ---- scratch carve-up
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    size_t o_keys = carve(total_keys * 4), o_sorted = carve(total_keys * 4);
    size_t o_counts = carve((size_t)nb * 4), o_offsets = carve((size_t)nb * 4), o_cursor = carve((size_t)nb * 4);
    size_t o_ntasks = carve((size_t)nb * 4), o_toffs = carve((size_t)nb * 4);
    size_t o_bsums = carve(((size_t)nb / 1024 + 2) * 4), o_totals = carve(16);
    size_t o_tasks = carve(max_tasks * sizeof(MsmTask)), o_tout = carve(max_tasks * sizeof(Pt));
    size_t o_perm = carve(max_tasks * 4), o_bins = carve(3 * MSM_TASK_CAP * 4);
    // at most total_keys / CAP buckets can hold more than CAP points (the list is a superset bound)
    const uint32_t heavy_cap = (uint32_t)(total_keys / MSM_TASK_CAP + 1);
    size_t o_heavy = carve((size_t)(heavy_cap + 1) * 4);
    // reduce tree: level l holds W * ceil(M / L^(l+1)) (A, U) pairs; two ping-pong buffers of the level-0 size
    const uint32_t n_lvl0 = (M + MSM_RED_L - 1) >> MSM_RED_LOG_L;
    size_t o_red = carve((size_t)4 * WB * n_lvl0 * sizeof(Pt));
    size_t o_S = carve((size_t)WB * sizeof(Pt));
    void *base;
    ZKB_TRY(ctx_scratch(ctx, "msm", off, &base));
    char *B = (char *)base;
    uint32_t *keys = (uint32_t *)(B + o_keys), *sorted = (uint32_t *)(B + o_sorted);
    uint32_t *counts = (uint32_t *)(B + o_counts), *offsets = (uint32_t *)(B + o_offsets), *cursor = (uint32_t *)(B + o_cursor);
    uint32_t *ntasks = (uint32_t *)(B + o_ntasks), *toffs = (uint32_t *)(B + o_toffs);
    uint32_t *bsums = (uint32_t *)(B + o_bsums), *totals = (uint32_t *)(B + o_totals);
    MsmTask *tasks = (MsmTask *)(B + o_tasks);
    uint32_t *n_heavy = (uint32_t *)(B + o_heavy), *heavy = n_heavy + 1;
    uint32_t *perm = (uint32_t *)(B + o_perm), *bins = (uint32_t *)(B + o_bins), *bin_off = bins + MSM_TASK_CAP,
             *bin_cur = bins + 2 * MSM_TASK_CAP;
    Pt *tout = (Pt *)(B + o_tout), *red = (Pt *)(B + o_red), *S = (Pt *)(B + o_S);

    // counts and cursor are adjacent carve-outs: one memset
    ZKB_CUDA_OK(ctx, cudaMemsetAsync(counts, 0, (size_t)nb * 4, st));
    ZKB_CUDA_OK(ctx, cudaMemsetAsync(cursor, 0, (size_t)nb * 4, st));
    msm_digits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((uint32_t)n, (const uint32_t *)d_scalars, c, W, tabled ? 0u : M, keys, counts);
    ZKB_CUDA_OK(ctx, cudaMemsetAsync(n_heavy, 0, 4, st));
    msm_ntasks_kernel<<<(nb + 255) / 256, 256, 0, st>>>(nb, counts, ntasks, n_heavy, heavy, heavy_cap);
    ctx->launches += 2;
    ZKB_TRY(exclusive_scan(ctx, nb, counts, offsets, bsums, totals, st));
    ZKB_TRY(exclusive_scan(ctx, nb, ntasks, toffs, bsums, totals + 1, st));
    msm_fill_tasks_kernel<<<(nb + 255) / 256, 256, 0, st>>>(nb, counts, offsets, toffs, tasks);
    msm_scatter_kernel<<<(unsigned)((total_keys + 255) / 256), 256, 0, st>>>(total_keys, (uint32_t)n, tabled ? (uint32_t)bases->n : 0u, keys, offsets, cursor, sorted);
    const Affine<F> *pts = (const Affine<F> *)(tabled ? bases->d_table : bases->d_points) + offset;
    const unsigned task_blocks = (unsigned)((max_tasks + 255) / 256);
    ZKB_CUDA_OK(ctx, cudaMemsetAsync(bins, 0, MSM_TASK_CAP * 4, st));
    msm_task_hist_kernel<<<task_blocks, 256, 0, st>>>(totals + 1, tasks, bins);
    msm_task_scan_kernel<<<1, MSM_TASK_CAP, 0, st>>>(bins, bin_off, bin_cur);
    msm_task_permute_kernel<<<task_blocks, 256, 0, st>>>(totals + 1, tasks, bin_off, bin_cur, perm);
    msm_accumulate_kernel<F><<<(unsigned)((max_tasks + 127) / 128), 128, 0, st>>>(totals + 1, tasks, perm, sorted, pts, tout);
    msm_heavy_tree_kernel<F><<<heavy_cap < 1024 ? heavy_cap : 1024, 256, 0, st>>>(n_heavy, heavy, heavy_cap, toffs, ntasks, tout);
    ctx->launches += 7;
    {
        uint32_t n_in = M;
        int span_bits = 0, level = 0;
        const size_t half = (size_t)2 * WB * n_lvl0;   // elements per ping-pong buffer (A then U)
        while (true) {
            uint32_t n_out = (n_in + MSM_RED_L - 1) >> MSM_RED_LOG_L;
            Pt *src = red + (size_t)((level + 1) & 1) * half, *dst = red + (size_t)(level & 1) * half;
            const Pt *inA = level == 0 ? tout : src, *inU = level == 0 ? nullptr : src + (size_t)WB * n_lvl0;
            const bool root = n_out == 1;
            uint32_t cnt = (uint32_t)WB * n_out;
            if (level == 0) {
                msm_reduce_level_kernel<F, 2><<<(cnt * 2 + 127) / 128, 128, 0, st>>>(WB, n_in, span_bits, root, toffs, ntasks, inA, inU,
                                                                                   root ? S : dst, dst + (size_t)WB * n_lvl0);
            } else if (F::N <= 12 && cnt <= 2048) {   // few nodes left: latency matters, quad-cooperative additions
                if constexpr (F::N <= 12)
                msm_reduce_level_quad_kernel<F><<<(cnt * 16 + 127) / 128, 128, 0, st>>>(WB, n_in, span_bits, root, inA, inU, root ? S : dst,
                                                                                      dst + (size_t)WB * n_lvl0);
            } else {
                msm_reduce_level_kernel<F, 4><<<(cnt * 4 + 127) / 128, 128, 0, st>>>(WB, n_in, span_bits, root, toffs, ntasks, inA, inU,
                                                                                   root ? S : dst, dst + (size_t)WB * n_lvl0);
            }
            ctx->launches++;
            if (root) break;
            n_in = n_out;
            span_bits += MSM_RED_LOG_L;
            level++;
        }
    }
    ZKB_CUDA_OK(ctx, cudaGetLastError());
    std::vector<uint32_t> hs((size_t)WB * 4 * F::N);
    ZKB_CUDA_OK(ctx, cudaMemcpyAsync(hs.data(), S, hs.size() * 4, cudaMemcpyDeviceToHost, st));
    ZKB_CUDA_OK(ctx, cudaStreamSynchronize(st));
    return msm_window_combine(bases->curve, c, WB, hs.data(), partial_host);
}

static int msm_run(zkb_ctx *ctx, const zkb_msm_bases *bases, uint64_t offset, uint64_t n, const void *d_scalars,
                   uint32_t *partial_host, cudaStream_t st) {
    ZKB_DISPATCH_CURVE(bases->curve, return msm_run_t<CF, SB>(ctx, bases, offset, n, d_scalars, partial_host, st))
    return ZKB_ERR_INVALID_ARGUMENT;
}

// u32 limbs per affine coordinate (x or y)
static int curve_coord_limbs(int curve) {
    ZKB_DISPATCH_CURVE(curve, return CF::N)
    return 0;
}

// ------------------------------------------------------------------------------------ synthetic points
// out[i] = A[i % m] + B[i / m] (affine, canonical): lets benchmarks and tests build millions of valid,
// distinct curve points from two small host-made tables ("generate on GPU").
template <class F>
__global__ void __launch_bounds__(128) grid_points_kernel(uint64_t n, uint32_t m, const Affine<F> *__restrict__ A,
                                                          const Affine<F> *__restrict__ Bt, Affine<F> *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    XYZZ<F> acc = XYZZ<F>::from_affine(A[i % m].to_mont());
    acc.add_mixed(Bt[i / m].to_mont());
    out[i] = acc.to_affine().from_mont();
}

template <class F>
static int grid_points_t(zkb_ctx *ctx, uint64_t n, uint32_t m, const void *dA, const void *dB, void *dout, cudaStream_t st) {
    typedef Affine<F> A;
    grid_points_kernel<F><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, m, (const A *)dA, (const A *)dB, (A *)dout);
    ctx->launches++;
    ZKB_CUDA_OK(ctx, cudaGetLastError());
    return ZKB_OK;
}

// ------------------------------------------------------------------------------------ fixed-base batch exponentiation
// algebra::batch_exp<G, Fr>(scalar_size, window, table, v) / windowed_exp as the Groth16 generator calls them
// (r1cs_gg_ppzksnark/generator.hpp:167-225, knowledge_commitment_multiexp.hpp:110-205): out[i] = v_i * base for one
// base and many scalars.  Table: d * 2^(8 w) * base for the 32 byte windows of a 256-bit scalar, d = 1 .. 255 (one thread
// per entry: doublings to its window, double-and-add over the byte, one inversion - the table is affine so the main pass
// uses mixed additions).  Main pass: one thread per scalar, at most 32 mixed additions and one inversion to hand back
// affine points (the form every consumer here takes).  Setup-time operation: no bucket machinery.
#define BEXP_WINDOWS 32
template <class F>
__global__ void __launch_bounds__(128) batch_exp_table_kernel(Affine<F> base_mont, Affine<F> *__restrict__ table) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= BEXP_WINDOWS * 255) return;
    const uint32_t w = i / 255, d = i % 255 + 1;
    XYZZ<F> b = XYZZ<F>::from_affine(base_mont);
    for (uint32_t k = 0; k < 8 * w; k++) b = b.dbl();
    table[i] = b.mul_small(d).to_affine();
}

template <class F>
__global__ void __launch_bounds__(128) batch_exp_kernel(uint64_t n, const uint32_t *__restrict__ scalars,
                                                        const Affine<F> *__restrict__ table, Affine<F> *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    XYZZ<F> acc = XYZZ<F>::infinity();
    for (int w = 0; w < BEXP_WINDOWS; w++) {
        const uint32_t d = (scalars[i * 8 + w / 4] >> (8 * (w % 4))) & 0xffu;
        if (d) acc.add_mixed(table[w * 255 + d - 1]);
    }
    out[i] = acc.to_affine().from_mont();      // infinity -> the all-zero encoding
}

template <class F>
static int batch_exp_t(zkb_ctx *ctx, uint64_t n, const uint32_t *base_affine, const void *d_scalars, void *d_out, cudaStream_t st) {
    typedef Affine<F> A;
    A base;
    memcpy(&base, base_affine, sizeof(A));
    void *tab;
    ZKB_TRY(ctx_scratch(ctx, "bexp_table", (size_t)BEXP_WINDOWS * 255 * sizeof(A), &tab));
    batch_exp_table_kernel<F><<<(BEXP_WINDOWS * 255 + 127) / 128, 128, 0, st>>>(base.to_mont(), (A *)tab);
    batch_exp_kernel<F><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, (const uint32_t *)d_scalars, (const A *)tab, (A *)d_out);
    ctx->launches += 2;
    ZKB_CUDA_OK(ctx, cudaGetLastError());
    return ZKB_OK;
}

// ------------------------------------------------------------------------------------ window table
// table[w * n + i] = 2^(c w) P_i for w = 1 .. W-1 (affine, Montgomery form; row 0 is a copy of the bases).
// A commitment key / proving-key query vector is long-lived, so this is paid once: afterwards every digit
// window of an MSM adds into the same 2^(c-1) buckets (no per-window bucket sets, no window combine) and c
// can be as large as log2 n, which cuts the number of windows W = ceil((bits+1)/c).
template <class F>
__global__ void __launch_bounds__(128) msm_table_kernel(uint64_t n, int c, int W, const Affine<F> *__restrict__ pts,
                                                        Affine<F> *__restrict__ table) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine<F> p = pts[i];
    table[i] = p;
    XYZZ<F> acc = XYZZ<F>::from_affine(p);
    for (int w = 1; w < W; w++) {
        for (int k = 0; k < c; k++) acc = acc.dbl();
        p = acc.to_affine();                 // one inversion per point and window (setup cost)
        table[(uint64_t)w * n + i] = p;
        acc = XYZZ<F>::from_affine(p);       // keep ZZ = ZZZ = 1: the next c doublings start cheap
    }
}

template <class F, int SCALAR_BITS>
static int msm_table_t(zkb_ctx *ctx, zkb_msm_bases *b, int c, uint64_t max_bytes, cudaStream_t st) {
    const int W = (SCALAR_BITS + 1 + c - 1) / c;
    const uint64_t bytes = (uint64_t)W * b->n * sizeof(Affine<F>);
    if (bytes > max_bytes || (uint64_t)W * b->n >= (1ull << 31))
        return ctx_fail(ctx, ZKB_ERR_OUT_OF_MEMORY, "zkb_msm_bases_precompute: table exceeds max_bytes");
    void *t = nullptr;
    cudaError_t e = cudaMalloc(&t, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return ctx_fail(ctx, ZKB_ERR_OUT_OF_MEMORY, "cudaMalloc MSM window table");
    }
    msm_table_kernel<F><<<(unsigned)((b->n + 127) / 128), 128, 0, st>>>(b->n, c, W, (const Affine<F> *)b->d_points, (Affine<F> *)t);
    ctx->launches++;
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) {
        cudaFree(t);
        return ctx_fail(ctx, ZKB_ERR_CUDA, std::string("msm_table_kernel: ") + cudaGetErrorString(e));
    }
    b->d_table = t;
    b->table_c = c;
    b->table_w = W;
    return ZKB_OK;
}

// ------------------------------------------------------------------------------------ C ABI
extern "C" {

int zkb_batch_exp(zkb_ctx *ctx, int curve, uint64_t n, const uint32_t *base_affine, const void *scalars, void *out_affine, int mem,
                  void *stream) {
    if (!ctx) return ZKB_ERR_INVALID_ARGUMENT;
    int cl = curve_coord_limbs(curve);
    if (!cl || !base_affine || (n && (!scalars || !out_affine))) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_batch_exp: bad arguments");
    if (n == 0) return ZKB_OK;
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t pb = (size_t)2 * cl * 4;
    const void *d_sc = scalars;
    void *d_out = out_affine;
    if (mem != ZKB_MEM_DEVICE) {
        void *p, *q;
        ZKB_TRY(ctx_scratch(ctx, "io_in", n * 32, &p));
        ZKB_TRY(ctx_scratch(ctx, "io_out", n * pb, &q));
        ZKB_CUDA_OK(ctx, cudaMemcpyAsync(p, scalars, n * 32, cudaMemcpyHostToDevice, st));
        d_sc = p;
        d_out = q;
    }
    int s = ZKB_ERR_INVALID_ARGUMENT;
    ZKB_DISPATCH_CURVE(curve, s = batch_exp_t<CF>(ctx, n, base_affine, d_sc, d_out, st))
    ZKB_TRY(s);
    if (mem != ZKB_MEM_DEVICE) {
        ZKB_CUDA_OK(ctx, cudaMemcpyAsync(out_affine, d_out, n * pb, cudaMemcpyDeviceToHost, st));
        ZKB_CUDA_OK(ctx, cudaStreamSynchronize(st));
    }
    return ZKB_OK;
}

int zkb_g1_grid_points(zkb_ctx *ctx, int curve, uint64_t n, uint32_t m, const void *table_a, const void *table_b,
                       void *out_device, void *stream) {
    if (!ctx) return ZKB_ERR_INVALID_ARGUMENT;
    int cl = curve_coord_limbs(curve);
    if (!cl || !m || !table_a || !table_b || (n && !out_device))
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_g1_grid_points: bad arguments");
    if (n == 0) return ZKB_OK;
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    size_t pb = (size_t)2 * cl * 4;
    uint64_t nb = (n + m - 1) / m;
    void *dA, *dB;
    ZKB_TRY(ctx_scratch(ctx, "grid_a", m * pb, &dA));
    ZKB_TRY(ctx_scratch(ctx, "grid_b", nb * pb, &dB));
    ZKB_CUDA_OK(ctx, cudaMemcpyAsync(dA, table_a, m * pb, cudaMemcpyHostToDevice, st));
    ZKB_CUDA_OK(ctx, cudaMemcpyAsync(dB, table_b, nb * pb, cudaMemcpyHostToDevice, st));
    int s = ZKB_ERR_INVALID_ARGUMENT;
    ZKB_DISPATCH_CURVE(curve, s = grid_points_t<CF>(ctx, n, m, dA, dB, out_device, st))
    ZKB_TRY(s);
    ZKB_CUDA_OK(ctx, cudaStreamSynchronize(st));
    return ZKB_OK;
}

int zkb_msm_bases_create(zkb_ctx *ctx, int curve, uint64_t n, const void *points_affine, int mem, void *stream,
                         zkb_msm_bases **out) {
    if (!ctx || !out) return ZKB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int cl = curve_coord_limbs(curve);
    if (!cl || (n && !points_affine)) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_msm_bases_create: bad curve/pointer");
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    size_t bytes = (size_t)n * 2 * cl * 4;
    void *d = nullptr;
    if (n) {
        cudaError_t e = cudaMalloc(&d, bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return ctx_fail(ctx, ZKB_ERR_OUT_OF_MEMORY, "cudaMalloc MSM bases");
        }
        cudaError_t ce = cudaMemcpyAsync(d, points_affine, bytes, mem == ZKB_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st);
        if (ce != cudaSuccess) {
            cudaFree(d);
            return ctx_fail(ctx, ZKB_ERR_CUDA, std::string("copy MSM bases: ") + cudaGetErrorString(ce));
        }
        unsigned blocks = (unsigned)((n + 255) / 256);
        ZKB_DISPATCH_CURVE(curve, points_to_mont_kernel<CF><<<blocks, 256, 0, st>>>(n, (Affine<CF> *)d))
        ctx->launches++;
        cudaError_t ke = cudaStreamSynchronize(st);
        if (ke != cudaSuccess) {
            cudaFree(d);
            return ctx_fail(ctx, ZKB_ERR_CUDA, std::string("points_to_mont: ") + cudaGetErrorString(ke));
        }
    }
    zkb_msm_bases *b = new zkb_msm_bases();
    b->device = ctx->device;
    b->curve = curve;
    b->n = n;
    b->d_points = d;
    *out = b;
    return ZKB_OK;
}

void zkb_msm_bases_free(zkb_msm_bases *b) {
    if (!b) return;
    if (b->d_points || b->d_table) {
        cudaSetDevice(b->device);
        if (b->d_points) cudaFree(b->d_points);
        if (b->d_table) cudaFree(b->d_table);
    }
    delete b;
}

uint64_t zkb_msm_bases_size(const zkb_msm_bases *b) { return b ? b->n : 0; }

int zkb_msm_bases_precompute(zkb_ctx *ctx, zkb_msm_bases *b, int window_bits, uint64_t max_bytes, void *stream) {
    if (!ctx || !b || b->device != ctx->device) return ZKB_ERR_INVALID_ARGUMENT;
    if (b->d_table || b->n == 0) return ZKB_OK;
    int c = window_bits;
    if (c == 0) {   // one bucket per ~2 points of a full-length MSM
        c = 1;
        while ((1ull << c) < b->n) c++;
        if (c < 8) c = 8;
        if (c > 22) c = 22;
    }
    if (c < 2 || c > 24) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_msm_bases_precompute: window_bits out of range");
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    ZKB_DISPATCH_CURVE(b->curve, return msm_table_t<CF, SB>(ctx, b, c, max_bytes, st))
    return ZKB_ERR_INVALID_ARGUMENT;
}

int zkb_msm_partial(zkb_ctx *ctx, const zkb_msm_bases *bases, uint64_t offset, uint64_t n, const void *scalars, int mem,
                    uint32_t *partial_xyzz_host, void *stream) {
    if (!ctx || !bases || !partial_xyzz_host) return ZKB_ERR_INVALID_ARGUMENT;
    if (bases->device != ctx->device || offset > bases->n || n > bases->n - offset || (n && !scalars))
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_msm: range outside the bases / null scalars");
    int cl = curve_coord_limbs(bases->curve);
    if (n == 0) {  // empty sum = infinity (XYZZ all zero)
        memset(partial_xyzz_host, 0, (size_t)4 * cl * 4);
        return ZKB_OK;
    }
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const void *ds = scalars;
    if (mem != ZKB_MEM_DEVICE) {
        void *d;
        ZKB_TRY(ctx_scratch(ctx, "msm_scalars", (size_t)n * 32, &d));
        ZKB_CUDA_OK(ctx, cudaMemcpyAsync(d, scalars, (size_t)n * 32, cudaMemcpyHostToDevice, st));
        ds = d;
    }
    return msm_run(ctx, bases, offset, n, ds, partial_xyzz_host, st);
}

int zkb_msm(zkb_ctx *ctx, const zkb_msm_bases *bases, uint64_t offset, uint64_t n, const void *scalars, int mem,
            uint32_t *result_affine, void *stream) {
    if (!ctx || !bases || !result_affine) return ZKB_ERR_INVALID_ARGUMENT;
    uint32_t partial[4 * 24];
    ZKB_TRY(zkb_msm_partial(ctx, bases, offset, n, scalars, mem, partial, stream));
    return zkb_msm_combine(bases->curve, 1, partial, result_affine);
}

int zkb_msm_g1(zkb_ctx *ctx, int curve, uint64_t n, const void *points_affine, const void *scalars, int mem,
               uint32_t *result_affine, void *stream) {
    zkb_msm_bases *b = nullptr;
    ZKB_TRY(zkb_msm_bases_create(ctx, curve, n, points_affine, mem, stream, &b));
    int s = zkb_msm(ctx, b, 0, n, scalars, mem, result_affine, stream);
    zkb_msm_bases_free(b);
    return s;
}

}  // extern "C"
