// Placeholder argument builders on the device: expression evaluation over an extended domain, the
// quotient division + split, and the lookup argument's sort.  With these the columns of a Placeholder round stay in HBM
// between the commits instead of round-tripping through the host.
//
// Reference:
//   gates argument        zk/snark/systems/plonk/placeholder/gates_argument.hpp:76-217  (F = mask * sum_gates selector *
//                         sum_constraints theta^k constraint(columns, rotations), as polynomial_dfs on a domain of
//                         rows * 2^ceil(log2(degree + 1)) points)
//   permutation argument  .../permutation_argument.hpp:133-215 (F_0..F_2: the same kind of expression over columns,
//                         S_id / S_sigma, V_P and V_P(omega X))
//   quotient              .../placeholder/prover.hpp:220-283: T = (sum alpha_i F_i).coefficients() / Z, split into chunks of
//                         `rows` coefficients, each chunk from_coefficients() on the basic domain
//   lookup sort           .../lookup_argument.hpp:565-633 (sort_polynomials)
//
// Expression evaluation works coset by coset: the extended domain of size E = D n is the union of the cosets
// w_E^j H (H the n-subgroup), a column's values on one coset are one coset NTT of its coefficients (zkb_ntt with a shift),
// and a rotation by r rows is an index shift by r inside the coset.  zkb_expr_eval runs a postfix program once per point
// of a coset over [ncols][n] arrays and writes out[offset + i * stride]: with offset = j and stride = D the results of the
// D cosets interleave into the evaluation form on the size-E subgroup, which is what polynomial_dfs holds upstream.
#include <string.h>
#include <vector>
#include "zkb_field.cuh"
#include "zkb_internal.h"

using namespace zkb;

#define ZKB_DISPATCH_PLONK_FIELD(field, FN, ...)                                  \
    switch (field) {                                                             \
        case ZKB_FIELD_BLS12_381_FR: return FN<params::Bls12381Fr>(__VA_ARGS__); \
        case ZKB_FIELD_BN254_FR: return FN<params::Bn254Fr>(__VA_ARGS__);        \
        case ZKB_FIELD_PALLAS_FP: return FN<params::PallasFp>(__VA_ARGS__);      \
        case ZKB_FIELD_PALLAS_FQ: return FN<params::PallasFq>(__VA_ARGS__);      \
        default: return ZKB_ERR_INVALID_ARGUMENT;                                \
    }

template <class F, class P>
static bool plonk_canonical(const uint32_t *l) {
    for (int i = F::N - 1; i >= 0; i--) {
        if (l[i] < P::mod(i)) return true;
        if (l[i] > P::mod(i)) return false;
    }
    return false;
}

// ------------------------------------------------------------------------------------------------ expression evaluation
#define EXPR_STACK 16

template <class P>
__global__ void __launch_bounds__(128) expr_eval_kernel(uint64_t n, const Fp<P> *__restrict__ cols, const zkb_expr_instr *__restrict__ prog,
                                                        uint32_t n_instr, const Fp<P> *__restrict__ consts, Fp<P> *__restrict__ out,
                                                        uint64_t out_stride, uint64_t out_offset, int accumulate) {
    typedef Fp<P> F;
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F stack[EXPR_STACK];
    int sp = 0;
    for (uint32_t k = 0; k < n_instr; k++) {
        const zkb_expr_instr in = prog[k];
        switch (in.op) {
            case ZKB_EXPR_PUSH_COL: {
                const uint64_t row = (i + (uint64_t)(int64_t)in.b) & (n - 1);     // n is a power of two: wraps both ways
                stack[sp++] = cols[(uint64_t)in.a * n + row].to_mont();
                break;
            }
            case ZKB_EXPR_PUSH_CONST: stack[sp++] = consts[in.a]; break;
            case ZKB_EXPR_ADD: sp--; stack[sp - 1] = stack[sp - 1] + stack[sp]; break;
            case ZKB_EXPR_SUB: sp--; stack[sp - 1] = stack[sp - 1] - stack[sp]; break;
            case ZKB_EXPR_MUL: sp--; stack[sp - 1] = stack[sp - 1] * stack[sp]; break;
            default: stack[sp - 1] = stack[sp - 1].neg(); break;   // ZKB_EXPR_NEG
        }
    }
    F r = stack[0].from_mont();
    F *dst = out + out_offset + i * out_stride;
    if (accumulate) r = r + *dst;
    *dst = r;
}

template <class P>
static int expr_eval_t(zkb_ctx *ctx, uint64_t n, uint32_t ncols, const void *d_cols, const zkb_expr_instr *program, uint32_t n_instr,
                       const uint32_t *constants, uint32_t nconst, void *d_out, uint64_t out_stride, uint64_t out_offset,
                       int accumulate, cudaStream_t st) {
    typedef Fp<P> F;
    // validate on the host: the kernel trusts the program
    int depth = 0;
    for (uint32_t k = 0; k < n_instr; k++) {
        const zkb_expr_instr &in = program[k];
        switch (in.op) {
            case ZKB_EXPR_PUSH_COL:
                if (in.a >= ncols) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_expr_eval: column index out of range");
                depth++;
                break;
            case ZKB_EXPR_PUSH_CONST:
                if (in.a >= nconst) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_expr_eval: constant index out of range");
                depth++;
                break;
            case ZKB_EXPR_ADD: case ZKB_EXPR_SUB: case ZKB_EXPR_MUL:
                if (depth < 2) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_expr_eval: stack underflow");
                depth--;
                break;
            case ZKB_EXPR_NEG:
                if (depth < 1) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_expr_eval: stack underflow");
                break;
            default: return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_expr_eval: unknown opcode");
        }
        if (depth > EXPR_STACK) return ctx_fail(ctx, ZKB_ERR_UNSUPPORTED, "zkb_expr_eval: expression needs more than 16 stack slots");
    }
    if (depth != 1) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_expr_eval: the program must leave exactly one value");
    std::vector<F> cm(nconst ? nconst : 1);
    for (uint32_t k = 0; k < nconst; k++) {
        if (!plonk_canonical<F, P>(constants + (size_t)k * F::N)) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_expr_eval: constant >= modulus");
        F c;
        memcpy(c.l, constants + (size_t)k * F::N, sizeof(c.l));
        cm[k] = c.to_mont();
    }
    const size_t pb = ((size_t)n_instr * sizeof(zkb_expr_instr) + 255) & ~(size_t)255, cb = cm.size() * sizeof(F);
    void *p;
    ZKB_TRY(ctx_scratch(ctx, "expr_prog", pb + cb, &p));
    ZKB_CUDA_OK(ctx, cudaMemcpyAsync(p, program, (size_t)n_instr * sizeof(zkb_expr_instr), cudaMemcpyHostToDevice, st));
    ZKB_CUDA_OK(ctx, cudaMemcpyAsync((char *)p + pb, cm.data(), cb, cudaMemcpyHostToDevice, st));
    ZKB_CUDA_OK(ctx, cudaStreamSynchronize(st));     // program / constants are host temporaries
    expr_eval_kernel<P><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(n, (const F *)d_cols, (const zkb_expr_instr *)p, n_instr,
                                                                     (const F *)((char *)p + pb), (F *)d_out, out_stride, out_offset, accumulate);
    ctx->launches++;
    ZKB_CUDA_OK(ctx, cudaGetLastError());
    return ZKB_OK;
}

// ------------------------------------------------------------------------------------------------ quotient
// t[k] = sum_{j >= 1, k + j n < E} c[k + j n]: the quotient of c(X) by X^n - 1 (remainder dropped), k < chunks * n
template <class P>
__global__ void __launch_bounds__(256) quotient_div_kernel(uint64_t n, uint64_t E, uint64_t count, const Fp<P> *__restrict__ c,
                                                           Fp<P> *__restrict__ t) {
    typedef Fp<P> F;
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= count) return;
    F acc = F::zero();
    for (uint64_t j = k + n; j < E; j += n) acc = acc + c[j];
    t[k] = acc;
}

template <class P>
static int quotient_split_t(zkb_ctx *ctx, int field, int log_n, int log_ext, const void *d_f, uint32_t nchunks, void *d_out, cudaStream_t st) {
    typedef Fp<P> F;
    const uint64_t n = 1ull << log_n, E = 1ull << log_ext, count = (uint64_t)nchunks * n;
    quotient_div_kernel<P><<<(unsigned)((count + 255) / 256), 256, 0, st>>>(n, E, count, (const F *)d_f, (F *)d_out);
    ctx->launches++;
    ZKB_CUDA_OK(ctx, cudaGetLastError());
    // from_coefficients on the basic domain, every chunk (prover.hpp:255-257)
    return ntt_device(ctx, field, log_n, nchunks, d_out, d_out, 0, nullptr, n, n, st);
}

// ------------------------------------------------------------------------------------------------ lookup sort
// sort_polynomials (lookup_argument.hpp:565-633).  The reference counts every value in an unordered_map (table values
// and lookup inputs), then walks the table values in order and, whenever the value changes, emits the previous value
// `count` times - a zero is emitted once (the initial `prev` is zero, and zero is the padding value), and a trailing
// zero run is not emitted at all.  Here: an open-addressing hash table keyed by the 256-bit value does the counting
// (slots hold the index of a representative element), a scan over the run heads gives every run its output position, and
// an expansion pass fills the output by binary search over those positions.
#define LS_EMPTY 0xffffffffu

__device__ __forceinline__ bool ls_equal(const uint4 *a, const uint4 *b) {
    const uint4 a0 = a[0], a1 = a[1], b0 = b[0], b1 = b[1];
    return a0.x == b0.x && a0.y == b0.y && a0.z == b0.z && a0.w == b0.w && a1.x == b1.x && a1.y == b1.y && a1.z == b1.z && a1.w == b1.w;
}
__device__ __forceinline__ bool ls_is_zero(const uint4 *a) {
    const uint4 a0 = a[0], a1 = a[1];
    return (a0.x | a0.y | a0.z | a0.w | a1.x | a1.y | a1.z | a1.w) == 0;
}
__device__ __forceinline__ uint32_t ls_hash(const uint4 *a, uint32_t mask) {
    const uint4 a0 = a[0], a1 = a[1];
    uint32_t h = a0.x * 0x9e3779b1u ^ a0.y * 0x85ebca77u ^ a0.z * 0xc2b2ae3du ^ a0.w * 0x27d4eb2fu ^ a1.x * 0x165667b1u ^ a1.y * 0xd3a2646cu ^
                 a1.z * 0xfd7046c5u ^ a1.w * 0xb55a4f09u;
    h ^= h >> 15;
    return h & mask;
}
// element e of the virtual concatenation: table values first ([n_values][usable] of the [n_values][n] array), then inputs
__device__ __forceinline__ const uint4 *ls_elem(uint64_t e, uint64_t n, uint64_t usable, uint64_t n_table, const uint4 *values, const uint4 *inputs) {
    if (e < n_table) return values + 2 * ((e / usable) * n + e % usable);
    e -= n_table;
    return inputs + 2 * ((e / usable) * n + e % usable);
}

// two passes: the table values insert (first = 0, count = n_table, insert = 1), then the lookup inputs only probe
// (first = n_table, insert = 0) - an input that reaches an empty slot is not in the table (the reference asserts, :577)
__global__ void __launch_bounds__(256) ls_count_kernel(uint64_t first, uint64_t count, int insert, uint64_t n, uint64_t usable, uint64_t n_table,
                                                       const uint4 *__restrict__ values, const uint4 *__restrict__ inputs,
                                                       uint32_t mask, uint32_t *slots, uint32_t *counts, uint32_t *missing) {
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= count) return;
    const uint64_t e = first + t;
    const uint4 *me = ls_elem(e, n, usable, n_table, values, inputs);
    uint32_t h = ls_hash(me, mask);
    while (true) {
        uint32_t cur = insert ? atomicCAS(slots + h, LS_EMPTY, (uint32_t)e) : slots[h];
        if (cur == LS_EMPTY) {
            if (!insert) { *missing = 1; return; }
            atomicAdd(counts + h, 1u);
            return;
        }
        if (ls_equal(me, ls_elem(cur, n, usable, n_table, values, inputs))) {
            atomicAdd(counts + h, 1u);
            return;
        }
        h = (h + 1) & mask;
    }
}

// emit[e] for every table position e: 0 unless e heads a run; a head emits what the reference emits when the run ENDS
// (count copies of a non-zero value, one copy of zero unless the run is the last one).  Position 0 also accounts for the
// virtual leading zero (prev = 0 before the walk).
__global__ void __launch_bounds__(256) ls_runs_kernel(uint64_t n_table, uint64_t n, uint64_t usable, const uint4 *__restrict__ values,
                                                      const uint4 *__restrict__ inputs, uint32_t mask, const uint32_t *__restrict__ slots, const uint32_t *__restrict__ counts,
                                                      uint32_t *__restrict__ emit, uint32_t *__restrict__ lead_zero, uint32_t *__restrict__ last_nonzero) {
    const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_table) return;
    const uint4 *me = ls_elem(e, n, usable, n_table, values, values);
    const bool zero = ls_is_zero(me);
    if (!zero) atomicMax(last_nonzero, (uint32_t)e + 1);      // 1 + the last non-zero position; 0 = the table is all zeros
    bool head;
    if (e == 0) {
        head = !zero;               // a leading zero run merges with the virtual zero, which heads it
        *lead_zero = 1;             // the virtual zero is emitted once when the first non-zero value arrives ..
    } else {
        head = !ls_equal(me, ls_elem(e - 1, n, usable, n_table, values, values));
    }
    uint32_t out = 0;
    if (head) {
        if (zero) {
            // one zero, unless no value follows this run (the reference never flushes a trailing zero run)
            out = 1;   // corrected by ls_trailing_zero_kernel when the run reaches the end
        } else {
            uint32_t h = ls_hash(me, mask);
            while (!ls_equal(me, ls_elem(slots[h], n, usable, n_table, values, inputs))) h = (h + 1) & mask;
            out = counts[h];
        }
    }
    emit[e] = out;
}

// one thread: if the table ends in a zero run, that run (or the virtual leading zero when the whole table is zero)
// emits nothing
__global__ void ls_trailing_zero_kernel(uint64_t n_table, uint32_t *__restrict__ emit, uint32_t *__restrict__ lead_zero,
                                        const uint32_t *__restrict__ last_nonzero) {
    const uint64_t start = *last_nonzero;            // first position of the trailing zero run
    if (start == n_table) return;                    // the table ends in a non-zero value
    if (start == 0) *lead_zero = 0;                  // all zeros (or empty): the virtual zero is never flushed
    else emit[start] = 0;
}

// ---- exclusive scan over uint32 (block of 1024 items per 256 threads, three launches)
__global__ void __launch_bounds__(256) ls_scan_block_kernel(uint64_t n, const uint32_t *__restrict__ in, uint32_t *__restrict__ out,
                                                            uint32_t *__restrict__ block_sums) {
    __shared__ uint32_t warp_sums[8];
    const uint64_t base = (uint64_t)blockIdx.x * 1024 + threadIdx.x * 4;
    uint32_t v[4], s = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
        v[k] = base + k < n ? in[base + k] : 0;
        s += v[k];
    }
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
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
__global__ void ls_scan_sums_kernel(uint32_t nblocks, uint32_t *block_sums, uint32_t *total) {   // one thread: <= a few thousand blocks
    uint32_t run = 0;
    for (uint32_t i = 0; i < nblocks; i++) {
        uint32_t v = block_sums[i];
        block_sums[i] = run;
        run += v;
    }
    *total = run;
}
__global__ void __launch_bounds__(256) ls_scan_add_kernel(uint64_t n, uint32_t *__restrict__ out, const uint32_t *__restrict__ block_sums) {
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] += block_sums[i >> 10];
}

// sorted[q / usable][q % usable] for output position q: the leading virtual zero first, then the runs in table order
__global__ void __launch_bounds__(256) ls_expand_kernel(uint64_t n_table, uint64_t n, uint64_t usable, uint64_t capacity,
                                                        const uint4 *__restrict__ values, const uint32_t *__restrict__ offs,
                                                        const uint32_t *__restrict__ total, const uint32_t *__restrict__ lead_zero,
                                                        uint4 *__restrict__ sorted, uint32_t *__restrict__ overflow) {
    const uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const uint64_t lz = *lead_zero, emitted = lz + *total;
    if (q == 0 && emitted > capacity) *overflow = 1;
    if (q >= emitted || q >= capacity) return;
    uint4 v0 = make_uint4(0, 0, 0, 0), v1 = v0;
    if (q >= lz) {
        // the run whose [offs, offs + emit) holds q - lz: the last table position with offs <= q - lz (later positions with
        // the same offset emit nothing, and the search lands on the one that does because it is the last with offs <= target
        // among those ... positions with emit 0 share the offset of the next head, so take the LAST index with offs <= target)
        const uint32_t target = (uint32_t)(q - lz);
        uint64_t lo = 0, hi = n_table;   // invariant: offs[lo] <= target, answer in [lo, hi)
        while (hi - lo > 1) {
            const uint64_t mid = (lo + hi) >> 1;
            if (offs[mid] <= target) lo = mid; else hi = mid;
        }
        const uint4 *src = ls_elem(lo, n, usable, n_table, values, values);
        v0 = src[0];
        v1 = src[1];
    }
    uint4 *dst = sorted + 2 * ((q / usable) * n + q % usable);
    dst[0] = v0;
    dst[1] = v1;
}
// sorted[i][usable] = sorted[i + 1][0] (lookup_argument.hpp:630-632)
__global__ void ls_link_kernel(uint32_t ncols, uint64_t n, uint64_t usable, uint4 *sorted) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i + 1 >= ncols || usable >= n) return;
    sorted[2 * ((uint64_t)i * n + usable)] = sorted[2 * ((uint64_t)(i + 1) * n)];
    sorted[2 * ((uint64_t)i * n + usable) + 1] = sorted[2 * ((uint64_t)(i + 1) * n) + 1];
}

extern "C" {

int zkb_expr_eval(zkb_ctx *ctx, int field, uint64_t n, uint32_t ncols, const void *cols_device, const zkb_expr_instr *program,
                  uint32_t n_instr, const uint32_t *constants, uint32_t nconst, void *out_device, uint64_t out_stride,
                  uint64_t out_offset, int accumulate, void *stream) {
    if (!ctx) return ZKB_ERR_INVALID_ARGUMENT;
    if (!program || n_instr == 0 || !out_device || (ncols && !cols_device) || (nconst && !constants) || n == 0 || (n & (n - 1)) || out_stride == 0)
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_expr_eval: bad arguments (n must be a power of two)");
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    ZKB_DISPATCH_PLONK_FIELD(field, expr_eval_t, ctx, n, ncols, cols_device, program, n_instr, constants, nconst, out_device, out_stride,
                             out_offset, accumulate, (cudaStream_t)stream)
}

int zkb_quotient_split(zkb_ctx *ctx, int field, int log_n, int log_ext, const void *f_coefficients_device, uint32_t nchunks,
                       void *out_dfs_device, void *stream) {
    if (!ctx) return ZKB_ERR_INVALID_ARGUMENT;
    if (!f_coefficients_device || !out_dfs_device || nchunks == 0 || log_n < 0 || log_ext < log_n || log_ext > 40)
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_quotient_split: bad arguments");
    if (log_n > zkb_field_two_adicity(field)) return ctx_fail(ctx, ZKB_ERR_DOMAIN_TOO_LARGE, "2^log_n exceeds the two-adicity");
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    ZKB_DISPATCH_PLONK_FIELD(field, quotient_split_t, ctx, field, log_n, log_ext, f_coefficients_device, nchunks, out_dfs_device, (cudaStream_t)stream)
}

int zkb_lookup_sort(zkb_ctx *ctx, int field, uint64_t n, uint64_t usable_rows, uint32_t n_inputs, const void *inputs_device,
                    uint32_t n_values, const void *values_device, void *sorted_device, void *stream) {
    if (!ctx) return ZKB_ERR_INVALID_ARGUMENT;
    if (zkb_field_limbs(field) != 8 || n == 0 || usable_rows == 0 || usable_rows > n || n_values == 0 || !values_device || !sorted_device ||
        (n_inputs && !inputs_device))
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_lookup_sort: bad arguments");
    const uint64_t n_table = (uint64_t)n_values * usable_rows, total = n_table + (uint64_t)n_inputs * usable_rows;
    const uint32_t ncols = n_inputs + n_values;
    if (total >= (1ull << 31)) return ctx_fail(ctx, ZKB_ERR_UNSUPPORTED, "zkb_lookup_sort: more than 2^31 values");
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    uint32_t cap = 1024;
    while (cap < 2 * total) cap <<= 1;
    const uint32_t nblk = (uint32_t)((n_table + 1023) / 1024);
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_slots = carve((size_t)cap * 4), o_counts = carve((size_t)cap * 4), o_emit = carve(n_table * 4), o_offs = carve(n_table * 4),
                 o_bs = carve((size_t)(nblk + 1) * 4), o_misc = carve(32);
    void *base;
    ZKB_TRY(ctx_scratch(ctx, "lookup_sort", off, &base));
    char *B = (char *)base;
    uint32_t *slots = (uint32_t *)(B + o_slots), *counts = (uint32_t *)(B + o_counts), *emit = (uint32_t *)(B + o_emit),
             *offs = (uint32_t *)(B + o_offs), *bs = (uint32_t *)(B + o_bs), *misc = (uint32_t *)(B + o_misc);   // total, lead_zero, overflow, missing, 1 + last non-zero table position
    ZKB_CUDA_OK(ctx, cudaMemsetAsync(slots, 0xff, (size_t)cap * 4, st));
    ZKB_CUDA_OK(ctx, cudaMemsetAsync(counts, 0, (size_t)cap * 4, st));
    ZKB_CUDA_OK(ctx, cudaMemsetAsync(misc, 0, 32, st));
    ZKB_CUDA_OK(ctx, cudaMemsetAsync(sorted_device, 0, (size_t)ncols * n * 32, st));
    const uint4 *vals = (const uint4 *)values_device, *inps = (const uint4 *)inputs_device;
    ls_count_kernel<<<(unsigned)((n_table + 255) / 256), 256, 0, st>>>(0, n_table, 1, n, usable_rows, n_table, vals, inps, cap - 1, slots, counts, misc + 3);
    if (total > n_table)
        ls_count_kernel<<<(unsigned)((total - n_table + 255) / 256), 256, 0, st>>>(n_table, total - n_table, 0, n, usable_rows, n_table, vals, inps, cap - 1, slots,
                                                                                  counts, misc + 3);
    ls_runs_kernel<<<(unsigned)((n_table + 255) / 256), 256, 0, st>>>(n_table, n, usable_rows, vals, inps, cap - 1, slots, counts, emit, misc + 1, misc + 4);
    ls_trailing_zero_kernel<<<1, 1, 0, st>>>(n_table, emit, misc + 1, misc + 4);
    ls_scan_block_kernel<<<nblk, 256, 0, st>>>(n_table, emit, offs, bs);
    ls_scan_sums_kernel<<<1, 1, 0, st>>>(nblk, bs, misc);
    ls_scan_add_kernel<<<(unsigned)((n_table + 255) / 256), 256, 0, st>>>(n_table, offs, bs);
    const uint64_t capacity = (uint64_t)ncols * usable_rows;
    ls_expand_kernel<<<(unsigned)((capacity + 255) / 256), 256, 0, st>>>(n_table, n, usable_rows, capacity, vals, offs, misc, misc + 1,
                                                                         (uint4 *)sorted_device, misc + 2);
    ls_link_kernel<<<(ncols + 127) / 128, 128, 0, st>>>(ncols, n, usable_rows, (uint4 *)sorted_device);
    ctx->launches += total > n_table ? 9 : 8;
    ZKB_CUDA_OK(ctx, cudaGetLastError());
    uint32_t h[4];
    ZKB_CUDA_OK(ctx, cudaMemcpyAsync(h, misc, 16, cudaMemcpyDeviceToHost, st));
    ZKB_CUDA_OK(ctx, cudaStreamSynchronize(st));
    if (h[3]) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_lookup_sort: a lookup input is not in the table");
    if (h[2]) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_lookup_sort: the sorted values do not fit (equal table values are not adjacent)");
    return ZKB_OK;
}

}  // extern "C"
