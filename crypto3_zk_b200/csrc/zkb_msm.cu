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
//   6 reduce      per window S_w = sum_k (k+1) B_k.  Level 0: every node turns MSM_RED_L = 8 buckets into
//                 (A = plain sum, U = sum j B_j, j local) with three short running sums spread over adjacent lanes.
//                 Tail ("bit tree"): S_w = sum U_s + sum A_s + 8 sum_s s A_s, and sum_s s A_s = sum_b 2^b T_b with
//                 T_b = sum of the A_s whose index has bit b set.  A stage kernel folds 256 entries per block in shared
//                 memory, 8 binary levels of mutually independent additions (depth 8, not 8 x 9), and hands on the
//                 plain sum plus the 8 subset sums; plain rows (U, earlier T_b) are just summed.  Two or three stages
//                 shrink 2^16 nodes to one, and one warp finishes with a quad-cooperative Horner over the bits.
//                 (Round 1 ran radix-8 (A, U) levels all the way up: 6 latency-bound launches with growing span
//                 doublings, 1.14 ms at 2^19 buckets; the tail is now ~0.25 ms.)
//   7 combine     sum_w 2^(c w) S_w on the host (c W doublings; zkb_msm_host.cpp)
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#ifndef ZKB_MSM_MUL_INLINE
#define ZKB_MUL_OUTLINE 1
#endif
#include "zkb_curve.cuh"
#include "zkb_internal.h"

using namespace zkb;

namespace zkb {
int msm_window_combine(int curve, int W, int base, int rem, const uint32_t *sums, uint32_t *out);
}

// Digit windows.  total = scalar bits + 1 (the spare bit absorbs the last carry of the signed recoding) is cut into
// W = ceil(total / c) windows of ALMOST EQUAL width: the first `rem` windows have base + 1 bits, the others base bits.
// With windows at multiples of c the top window of a 255-bit scalar keeps 256 mod c bits - 4 bits for c = 18: eight
// buckets then hold an eighth of all points each and the MSM waits for them (2^18 points: 4.6 ms instead of 2.6).
struct MsmWindows {
    int W, base, rem;
    __host__ __device__ int width(int w) const { return base + (w < rem ? 1 : 0); }
    __host__ __device__ int pos(int w) const { return w * base + (w < rem ? w : rem); }
    __host__ __device__ int max_width() const { return base + (rem ? 1 : 0); }
};
static MsmWindows msm_windows(int total_bits, int W) {
    MsmWindows m;
    m.W = W;
    m.base = total_bits / W;
    m.rem = total_bits % W;
    return m;
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
// A scalar with a bit at or above `scalar_bits` (not a canonical element of the scalar field) would overflow the top
// window's bucket range: it is skipped and reported through *bad (the call then fails with ZKB_ERR_INVALID_ARGUMENT).
__global__ void __launch_bounds__(256) msm_digits_kernel(uint32_t n, const uint32_t *__restrict__ scalars, MsmWindows win,
                                                         int scalar_bits, uint32_t win_stride, uint32_t *__restrict__ keys,
                                                         uint32_t *__restrict__ counts, uint32_t *__restrict__ bad) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int W = win.W;
    const uint4 *sp = reinterpret_cast<const uint4 *>(scalars) + 2 * (uint64_t)i;
    uint4 lo = sp[0], hi = sp[1];
    uint32_t l[9] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w, 0};
    if (scalar_bits < 256 && (hi.w >> (scalar_bits - 224)) != 0) {
        atomicOr(bad, 1u);
        for (int w = 0; w < W; w++) keys[(uint64_t)w * n + i] = MSM_SENTINEL;
        return;
    }
    uint32_t carry = 0;
    for (int w = 0; w < W; w++) {
        const int c = win.width(w);
        const uint32_t half = 1u << (c - 1), mask = (1u << c) - 1;
        uint32_t bit = win.pos(w), limb = bit >> 5, sh = bit & 31;
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
__global__ void __launch_bounds__(256) msm_ntasks_kernel(uint32_t nb, uint32_t task_cap, const uint32_t *__restrict__ counts,
                                                         uint32_t *__restrict__ ntasks, uint32_t *__restrict__ n_heavy,
                                                         uint32_t *__restrict__ heavy, uint32_t heavy_cap) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    uint32_t nt = (counts[b] + task_cap - 1) / task_cap;
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
__global__ void __launch_bounds__(256) msm_fill_tasks_kernel(uint32_t nb, uint32_t task_cap, const uint32_t *__restrict__ counts,
                                                             const uint32_t *__restrict__ offsets,
                                                             const uint32_t *__restrict__ task_offsets, MsmTask *tasks) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= nb) return;
    uint32_t cnt = counts[b], off = offsets[b], t = task_offsets[b];
    for (uint32_t s = 0; s < cnt; s += task_cap, t++) {
        MsmTask k;
        k.start = off + s;
        k.len = cnt - s < task_cap ? cnt - s : task_cap;
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
// Level 0.  Node (w, s) folds the buckets s*L .. s*L+L-1 of window w (bucket value = sum of the bucket's task results):
//     A = sum_j B_j                       ("run": running sum from the top bucket down)
//     U = sum_j j B_j,  j local           ("acc" += run after every step but the last)
// so that S_w = sum_k (k+1) B_k = sum_s (U_s + A_s + L s A_s); when the window has a single node (`root`) S_w = U + A is
// written directly.  The two running sums of a node are two dependent chains of XYZZ additions; they sit on adjacent
// lanes (acc trails run by one step through shared memory), so the chain is L + 1 additions long instead of 2 L.
// Output layout: outA[w * out_stride + s], outU[w * out_stride + s].
template <class F>
__global__ void __launch_bounds__(128) msm_reduce_level0_kernel(int W, uint32_t n_in, int root, uint32_t out_stride,
                                                                const uint32_t *__restrict__ task_offsets,
                                                                const uint32_t *__restrict__ ntasks,
                                                                const XYZZ<F> *__restrict__ tout,
                                                                XYZZ<F> *__restrict__ outA, XYZZ<F> *__restrict__ outU) {
    typedef XYZZ<F> Pt;
    constexpr int ROLE_RUN = 0, ROLE_ACC = 1;
    constexpr int NODES = 128 / 2;
    __shared__ Pt sh_run[2][NODES];
    const uint32_t n_out = (n_in + MSM_RED_L - 1) >> MSM_RED_LOG_L;
    const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t node = gtid / 2, ln = threadIdx.x / 2;
    const int role = gtid % 2;
    const bool live = node < (uint32_t)W * n_out;
    const uint32_t w = live ? node / n_out : 0, s = live ? node % n_out : 0;
    const uint32_t j0 = s << MSM_RED_LOG_L;
    uint32_t j1 = j0 + MSM_RED_L;
    if (j1 > n_in) j1 = n_in;
    Pt x = Pt::infinity();
    for (uint32_t step = 0; step <= MSM_RED_L; step++) {
        // step k handles bucket j = j1-1-k (run) and folds the run value of step k-1 into acc
        const bool has_item = live && step < j1 - j0;
        const uint64_t idx = (uint64_t)w * n_in + (j1 - 1 - step);
        // both roles funnel into ONE call site of add() (a warp runs its roles in lockstep)
        const Pt *src = nullptr;
        uint32_t cnt = 0;
        if (role == ROLE_RUN && has_item) {
            src = tout + task_offsets[idx];
            cnt = ntasks[idx];
        } else if (role == ROLE_ACC && live && step >= 1 && step < j1 - j0) {
            src = &sh_run[(step - 1) & 1][ln];   // run after bucket j1-step, which is > j0
            cnt = 1;
        }
        for (uint32_t t = 0; t < cnt; t++) x.add(src[t]);
        if (role == ROLE_RUN && has_item) sh_run[step & 1][ln] = x;
        __syncwarp();
    }
    if (!live) return;
    const uint64_t o = (uint64_t)w * out_stride + s;
    if (role == ROLE_RUN) {
        if (!root) outA[o] = x;
    } else {
        if (root) {   // S_w = U + A
            x.add(sh_run[(j1 - j0 - 1) & 1][ln]);
            outA[o] = x;
        } else {
            outU[o] = x;
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
template <class B>
__device__ __forceinline__ Fp2<B> quad_from(const Fp2<B> &x, int src) {
    Fp2<B> r;
    r.c0 = quad_from(x.c0, src);
    r.c1 = quad_from(x.c1, src);
    return r;
}
template <class B>
__device__ __forceinline__ Fp2<B> quad_pick(int sub, const Fp2<B> &a0, const Fp2<B> &a1, const Fp2<B> &a2, const Fp2<B> &a3) {
    Fp2<B> r;
    r.c0 = quad_pick(sub, a0.c0, a1.c0, a2.c0, a3.c0);
    r.c1 = quad_pick(sub, a0.c1, a1.c1, a2.c1, a3.c1);
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

// ---- bit tree: one stage ---------------------------------------------------------------------------
// in:  [rows_total][n] points, rows_total = W * R, row r of a window: 0 = the A chain (weighted by the entry index),
//      1 .. R-1 = plain rows (U, and the T_b rows earlier stages produced).
// out: [W][R + 8][n_blk], n_blk = ceil(n / 256): rows 0 .. R-1 = the block sums of the input rows; rows R + b (b < 8) =
//      T_b of the A row = sum of the block's entries whose LOCAL index has bit b set.
// A block folds 256 entries with 8 binary levels in shared memory.  Before level l (nodes of 2^l entries, 2N = 256 >> l
// of them) the regions are A, T_0 .. T_(l-1), 2N slots each; level l adds sibling pairs in every region (N (l+1)
// mutually independent additions <= 128 threads) and opens T_l[n] = A[2n+1].  Work: 2 additions per entry; depth: 8.
template <class F>
__global__ void __launch_bounds__(128) msm_bittree_stage_kernel(uint32_t n, uint32_t R, const XYZZ<F> *__restrict__ in,
                                                                XYZZ<F> *__restrict__ out) {
    typedef XYZZ<F> Pt;
    extern __shared__ __align__(16) unsigned char msm_tree_smem[];
    Pt *X = reinterpret_cast<Pt *>(msm_tree_smem);   // 256 slots
    const uint32_t n_blk = gridDim.x, j = blockIdx.x;
    const uint32_t row = blockIdx.y % R, w = blockIdx.y / R;
    const bool need_t = row == 0;
    const uint32_t t = threadIdx.x;
    const Pt *src = in + (uint64_t)blockIdx.y * n + (uint64_t)j * 256;
    const uint32_t count = n - j * 256 < 256 ? n - j * 256 : 256;
    {
        Pt a = 2 * t < count ? src[2 * t] : Pt::infinity();
        const Pt b = 2 * t + 1 < count ? src[2 * t + 1] : Pt::infinity();
        a.add(b);
        X[t] = a;
        if (need_t) X[128 + t] = b;
    }
    __syncthreads();
#pragma unroll 1
    for (uint32_t l = 1; l < 8; l++) {
        const uint32_t log_n = 7 - l, N = 1u << log_n;
        const uint32_t jobs = need_t ? N * (l + 1) : N;
        const bool live = t < jobs;
        const uint32_t r = t >> log_n, k = t & (N - 1);
        Pt lhs = Pt::infinity(), rhs = Pt::infinity();
        if (live) {
            lhs = X[r * 2 * N + 2 * k];
            rhs = X[r * 2 * N + 2 * k + 1];
        }
        __syncthreads();
        if (live) {
            lhs.add(rhs);
            X[r * N + k] = lhs;
            if (need_t && r == 0) X[(l + 1) * N + k] = rhs;
        }
        __syncthreads();
    }
    Pt *dst = out + (uint64_t)w * (R + 8) * n_blk + j;
    if (t == 0) dst[(uint64_t)row * n_blk] = X[0];
    if (need_t && t >= 1 && t <= 8) dst[(uint64_t)(R + t - 1) * n_blk] = X[t];
}

// ---- bit tree: finish ---------------------------------------------------------------------------------
// in: [W][R] points (every row reduced to one value): row 0 = sum A, row 1 = sum U, row 2 + b = T_b.
// S_w = sum U + sum A + 2^shift * sum_{b < nbits} 2^b T_b.  The doublings are the critical path (a lone warp needs ~7 us
// per group operation even with the operation spread over a quad), so the sum is folded as a binary tree instead of a
// Horner chain: Y[i] += 2^s Y[i + s] for s = 1, 2, 4, .. - depth nbits + log2(nbits) operations instead of 2 nbits.
// One block per window, one quad per pair.
template <class F>
__global__ void __launch_bounds__(128) msm_bittree_horner_kernel(uint32_t R, int nbits, int shift, const XYZZ<F> *__restrict__ in,
                                                                 XYZZ<F> *__restrict__ S) {
    typedef XYZZ<F> Pt;
    __shared__ Pt Y[32];
    const Pt *rows = in + (uint64_t)blockIdx.x * R;
    const int t = threadIdx.x, sub = t & 3, q = t >> 2, warp = t >> 5;
    if (t < 32) Y[t] = t < nbits ? rows[2 + t] : Pt::infinity();
    __syncthreads();
    for (int s = 1; s < nbits && s < 32; s <<= 1) {
        const int npairs = 16 / s;
        if (warp * 8 < npairs) {            // warp-uniform: the quad operations shuffle across the whole warp
            const bool live = q < npairs;
            const int i = q * 2 * s;
            Pt a = live ? Y[i] : Pt::infinity(), b = live ? Y[i + s] : Pt::infinity();
            for (int k = 0; k < s; k++) b = quad_dbl(b, sub);
            a = quad_add(a, b, sub);
            if (live && sub == 0) Y[i] = a;
        }
        __syncthreads();
    }
    if (warp == 0) {
        Pt acc = Y[0];
        for (int i = 0; i < shift; i++) acc = quad_dbl(acc, sub);
        acc = quad_add(acc, rows[0], sub);
        acc = quad_add(acc, rows[1], sub);
        if (t == 0) S[blockIdx.x] = acc;
    }
}

// ------------------------------------------------------------------------------------ host driver
static int env_int(const char *name, int dflt) {
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

// Window width without a table (every window has its own bucket set).  Large MSMs balance additions (n W) against
// bucket-reduce work (W 2^(c-1) buckets, two full additions each): c = log2 n - 4.  Small ones are latency bound - the
// chain of additions per bucket is what the accumulate kernel waits for - and prefer fewer points per bucket
// (profiles/r2g_msm_tune.txt: 2^16 1.77 ms at c = 15 vs 2.8 ms at c = 12; 2^18 3.0 ms at c = 16 vs 4.1 ms at c = 14).
static int msm_pick_c(uint64_t n) {
    int lg = 0;
    while ((1ull << lg) < n) lg++;
    int c = lg >= 20 ? lg - 4 : lg >= 18 ? lg - 2 : lg - 1;
    if (c < 2) c = 2;
    if (c > 20) c = 20;
    return env_int("ZKB_MSM_C", c);   // tuning experiments only (profiles/quick_msm_tune.py)
}
// Default window width of a window table (one bucket set for all windows): about log2 n, a little more for small vectors
// (2^16: 1.32 ms at c = 18 vs 1.59 ms at c = 16; 2^18: 2.35 ms at c = 19 vs 2.55 ms at c = 18).
static int msm_pick_table_c(uint64_t n) {
    int lg = 1;
    while ((1ull << lg) < n) lg++;
    int c = lg <= 16 ? lg + 2 : lg <= 18 ? lg + 1 : lg;
    if (c < 8) c = 8;
    if (c > 22) c = 22;
    return c;
}

// Longest run of additions one thread performs in the accumulate kernel.  A task is a dependent chain (one mixed addition
// takes a lone warp ~15 us), so on a small MSM the cap bounds the kernel's tail; every extra task of a bucket costs one
// full addition in the level-0 reduce, so large MSMs (whose kernel runs for milliseconds anyway) keep the long cap.
// Measured (profiles/r2e_msm_tune.txt): 2^20 is fastest with 256, 2^16 / 2^18 with 32-64.
static uint32_t msm_pick_task_cap(uint64_t total_keys) {
    const uint32_t cap = total_keys >= (8u << 20) ? MSM_TASK_CAP : 64u;
    return (uint32_t)env_int("ZKB_MSM_CAP", (int)cap);
}

template <class F, int SCALAR_BITS>
static int msm_run_t(zkb_ctx *ctx, const zkb_msm_bases *bases, uint64_t offset, uint64_t n, const void *d_scalars,
                     uint32_t *partial_host, cudaStream_t st) {
    typedef XYZZ<F> Pt;
    // with a window table every digit window adds 2^(c w) P_i straight into ONE bucket set
    const bool tabled = bases->d_table != nullptr;
    const int c_req = tabled ? bases->table_c : msm_pick_c(n);
    const int W = tabled ? bases->table_w : (SCALAR_BITS + 1 + c_req - 1) / c_req;   // digit windows
    const MsmWindows win = msm_windows(SCALAR_BITS + 1, W);
    const int c = win.max_width();                      // widest window: 2^(c-1) buckets per set
    const int WB = tabled ? 1 : W;                      // bucket sets
    const uint32_t M = 1u << (c - 1);
    const uint32_t nb = (uint32_t)WB * M;
    if (tabled && (uint64_t)W * bases->n >= (1ull << 31))
        return ctx_fail(ctx, ZKB_ERR_UNSUPPORTED, "MSM: window table does not cover the scalar width");
    const uint64_t total_keys = (uint64_t)W * n;
    if (n >= (1ull << 31) || total_keys >= (1ull << 32))
        return ctx_fail(ctx, ZKB_ERR_UNSUPPORTED, "MSM: W*n must be < 2^32 (split the range across calls/GPUs)");
    const uint32_t task_cap = msm_pick_task_cap(total_keys);
    const uint64_t max_tasks = (uint64_t)nb + total_keys / task_cap + 1;

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
    const uint32_t heavy_cap = (uint32_t)(total_keys / task_cap + 1);
    size_t o_heavy = carve((size_t)(heavy_cap + 1) * 4);
    // reduce: level 0 leaves (A, U) per node, [WB][2][n_lvl0]; bit-tree stage k turns [WB][R][n] into [WB][R + 8][ceil(n/256)]
    const uint32_t n_lvl0 = (M + MSM_RED_L - 1) >> MSM_RED_LOG_L;
    size_t o_red = carve((size_t)2 * WB * n_lvl0 * sizeof(Pt));
    size_t red2_elems = 0;
    {
        uint32_t nn = n_lvl0, R = 2;
        while (nn > 1) {
            nn = (nn + 255) / 256;
            R += 8;
            red2_elems += (size_t)WB * R * nn;
        }
    }
    size_t o_red2 = carve((red2_elems + 1) * sizeof(Pt));
    size_t o_S = carve((size_t)WB * sizeof(Pt) + 16);
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
    Pt *tout = (Pt *)(B + o_tout), *red = (Pt *)(B + o_red), *red2 = (Pt *)(B + o_red2), *S = (Pt *)(B + o_S);
    uint32_t *bad = (uint32_t *)(B + o_S + (size_t)WB * sizeof(Pt));   // set by the digits kernel on a non-canonical scalar

    // counts and cursor are adjacent carve-outs: one memset
    ZKB_CUDA_OK(ctx, cudaMemsetAsync(counts, 0, (size_t)nb * 4, st));
    ZKB_CUDA_OK(ctx, cudaMemsetAsync(cursor, 0, (size_t)nb * 4, st));
    ZKB_CUDA_OK(ctx, cudaMemsetAsync(bad, 0, 4, st));
    msm_digits_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>((uint32_t)n, (const uint32_t *)d_scalars, win, SCALAR_BITS, tabled ? 0u : M, keys, counts, bad);
    ZKB_CUDA_OK(ctx, cudaMemsetAsync(n_heavy, 0, 4, st));
    msm_ntasks_kernel<<<(nb + 255) / 256, 256, 0, st>>>(nb, task_cap, counts, ntasks, n_heavy, heavy, heavy_cap);
    ctx->launches += 2;
    ZKB_TRY(exclusive_scan(ctx, nb, counts, offsets, bsums, totals, st));
    ZKB_TRY(exclusive_scan(ctx, nb, ntasks, toffs, bsums, totals + 1, st));
    msm_fill_tasks_kernel<<<(nb + 255) / 256, 256, 0, st>>>(nb, task_cap, counts, offsets, toffs, tasks);
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
        const bool root = n_lvl0 == 1;
        msm_reduce_level0_kernel<F><<<((uint32_t)WB * n_lvl0 * 2 + 127) / 128, 128, 0, st>>>(
            WB, M, root, root ? 1u : 2u * n_lvl0, toffs, ntasks, tout, root ? S : red, red + n_lvl0);
        ctx->launches++;
        if (!root) {
            // bit tree over the n_lvl0 nodes of every window: S_w = sum U + sum A + L * sum_s s A_s
            const size_t smem = 256 * sizeof(Pt);
            ZKB_CUDA_OK(ctx, cudaFuncSetAttribute(msm_bittree_stage_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            uint32_t nn = n_lvl0, R = 2;
            int nbits = 0;
            while ((1u << nbits) < n_lvl0) nbits++;
            const Pt *src = red;
            Pt *dst = red2;
            while (nn > 1) {
                const uint32_t nb2 = (nn + 255) / 256;
                msm_bittree_stage_kernel<F><<<dim3(nb2, (uint32_t)WB * R), 128, smem, st>>>(nn, R, src, dst);
                ctx->launches++;
                src = dst;
                dst += (size_t)WB * (R + 8) * nb2;
                R += 8;
                nn = nb2;
            }
            msm_bittree_horner_kernel<F><<<WB, 128, 0, st>>>(R, nbits, MSM_RED_LOG_L, src, S);
            ctx->launches++;
        }
    }
    ZKB_CUDA_OK(ctx, cudaGetLastError());
    std::vector<uint32_t> hs((size_t)WB * 4 * F::N + 4);
    ZKB_CUDA_OK(ctx, cudaMemcpyAsync(hs.data(), S, hs.size() * 4, cudaMemcpyDeviceToHost, st));
    ZKB_CUDA_OK(ctx, cudaStreamSynchronize(st));
    if (hs[(size_t)WB * 4 * F::N] != 0)
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "MSM: a scalar is not a canonical element of the scalar field (>= 2^bits)");
    return msm_window_combine(bases->curve, WB, win.base, win.rem, hs.data(), partial_host);
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
// table[w * n + i] = 2^pos(w) P_i for w = 1 .. W-1 (affine, Montgomery form; row 0 is a copy of the bases), pos(w) the
// bit position of digit window w (MsmWindows: W = ceil((bits+1)/c) windows of almost equal width).
// A commitment key / proving-key query vector is long-lived, so this is paid once: afterwards every digit
// window of an MSM adds into the same 2^(c-1) buckets (no per-window bucket sets, no window combine) and c
// can be as large as log2 n, which cuts the number of windows W = ceil((bits+1)/c).
template <class F>
__global__ void __launch_bounds__(128) msm_table_kernel(uint64_t n, MsmWindows win, const Affine<F> *__restrict__ pts,
                                                        Affine<F> *__restrict__ table) {
    const int W = win.W;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Affine<F> p = pts[i];
    table[i] = p;
    XYZZ<F> acc = XYZZ<F>::from_affine(p);
    for (int w = 1; w < W; w++) {
        for (int k = 0; k < win.width(w - 1); k++) acc = acc.dbl();
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
    msm_table_kernel<F><<<(unsigned)((b->n + 127) / 128), 128, 0, st>>>(b->n, msm_windows(SCALAR_BITS + 1, W), (const Affine<F> *)b->d_points, (Affine<F> *)t);
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
    if (c == 0) c = msm_pick_table_c(b->n);
    if (c < 2 || c > 24) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_msm_bases_precompute: window_bits out of range");
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    ZKB_DISPATCH_CURVE(b->curve, return msm_table_t<CF, SB>(ctx, b, c, max_bytes, st))
    return ZKB_ERR_INVALID_ARGUMENT;
}

int zkb_msm_window_plan(const zkb_msm_bases *bases, uint64_t n, int *window_bits, int *windows, int *bucket_sets) {
    if (!bases) return ZKB_ERR_INVALID_ARGUMENT;
    int sb = 0;
    ZKB_DISPATCH_CURVE(bases->curve, sb = SB)
    if (!sb) return ZKB_ERR_INVALID_ARGUMENT;
    const bool tabled = bases->d_table != nullptr;
    const int c_req = tabled ? bases->table_c : msm_pick_c(n);
    const int W = tabled ? bases->table_w : (sb + 1 + c_req - 1) / c_req;
    const MsmWindows win = msm_windows(sb + 1, W);
    if (window_bits) *window_bits = win.max_width();
    if (windows) *windows = W;
    if (bucket_sets) *bucket_sets = tabled ? 1 : W;
    return ZKB_OK;
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
