// Callers and neighbours of the NTT/Merkle kernels on the LPC/FRI and Groth16 paths (SURVEY 8(a) rows a4, a5,
// a10, a11 and the first "next" row 8(f)-1), all on device-resident data:
//   zkb_fri_commit_phase   - commit phase of zk::algorithms::proof_eval<FRI>     basic_fri.hpp:706-737
//   zkb_poly_evaluate      - polys_evaluator::eval_polys                          batched_commitment.hpp:176-190
//   zkb_poly_lincomb       - sum_j theta^k g_j - sum_j theta^k z_j               lpc.hpp:139-153, 163-176
//   zkb_poly_div_linear    - Q_normal / (X - point)                              lpc.hpp:154, 177
//   zkb_sparse_matvec      - cs.constraints[i].{a,b,c}.evaluate(assignment)      r1cs_to_qap.hpp:245-248, 289-291
// Data convention as everywhere in the library: elements canonical in HBM, constants in Montgomery form, so
// mont_mul(canonical, montgomery constant) is the canonical product.
#include <stdio.h>
#include <string.h>
#include <vector>
#include "zkb_field.cuh"
#include "zkb_internal.h"

using namespace zkb;

#define ZKB_POLY_THREADS 256
#define ZKB_POLY_MAXPTS 4

#define ZKB_DISPATCH_FR(field, FN, ...)                                          \
    switch (field) {                                                             \
        case ZKB_FIELD_BLS12_381_FR: return FN<params::Bls12381Fr>(__VA_ARGS__); \
        case ZKB_FIELD_BN254_FR: return FN<params::Bn254Fr>(__VA_ARGS__);        \
        case ZKB_FIELD_PALLAS_FP: return FN<params::PallasFp>(__VA_ARGS__);      \
        case ZKB_FIELD_PALLAS_FQ: return FN<params::PallasFq>(__VA_ARGS__);      \
        default: return ZKB_ERR_INVALID_ARGUMENT;                                \
    }

static bool is_fr(int f) { return f >= ZKB_FIELD_BLS12_381_FR && f <= ZKB_FIELD_PALLAS_FQ; }

template <class F, class P>
static bool canonical(const uint32_t *l) {
    for (int i = F::N - 1; i >= 0; i--) {
        if (l[i] < P::mod(i)) return true;
        if (l[i] > P::mod(i)) return false;
    }
    return false;
}

// block-wide sum of one field element per thread (all threads must call; result valid in thread 0)
template <class F>
__device__ F block_sum(F v, F *sh) {
    const uint32_t t = threadIdx.x;
    sh[t] = v;
    __syncthreads();
    for (uint32_t s = ZKB_POLY_THREADS / 2; s > 0; s >>= 1) {
        if (t < s) sh[t] = sh[t] + sh[t + s];
        __syncthreads();
    }
    F r = sh[0];
    __syncthreads();
    return r;
}

// ------------------------------------------------------------------------------------ evaluation
// Segment s of polynomial b: sum_{k in segment} c_k z^(k - seg_start) for up to ZKB_POLY_MAXPTS points at once.
// Thread t walks k = seg_start + t + 256 m downwards with Horner steps in z^256 (coalesced reads: the 256 threads
// of a block read one contiguous 8 KB run per step), then the lanes are combined with z^t.
template <class F>
struct EvalPoints {
    F z[ZKB_POLY_MAXPTS];      // Montgomery
    F z256[ZKB_POLY_MAXPTS];   // z^256, Montgomery
    int count;
};

template <class P>
__global__ void __launch_bounds__(ZKB_POLY_THREADS) poly_eval_segments_kernel(const Fp<P> *coeffs, uint64_t n, uint64_t poly_stride,
                                                                              uint64_t seg_len, uint32_t segs, EvalPoints<Fp<P>> pts,
                                                                              int scale_by_start, int paired, Fp<P> *partial) {
    typedef Fp<P> F;
    __shared__ F sh[ZKB_POLY_THREADS];
    const uint32_t seg = blockIdx.x, b = blockIdx.y, t = threadIdx.x;
    const F *c = coeffs + (uint64_t)b * poly_stride;
    const uint64_t start = (uint64_t)seg * seg_len;
    uint64_t end = start + seg_len;
    if (end > n) end = n;
    F acc[ZKB_POLY_MAXPTS];
#pragma unroll
    for (int p = 0; p < ZKB_POLY_MAXPTS; p++) acc[p] = F::zero();
    if (start + t < end) {
        uint64_t steps = (end - start - t + ZKB_POLY_THREADS - 1) / ZKB_POLY_THREADS;
        for (uint64_t m = steps; m-- > 0;) {
            F v = c[start + t + m * ZKB_POLY_THREADS];
#pragma unroll
            for (int p = 0; p < ZKB_POLY_MAXPTS; p++)
                if (p < pts.count) acc[p] = acc[p] * pts.z256[p] + v;
        }
    }
    // paired: also the value at -z.  A lane's coefficients k = start + t + 256 m all have the parity of t (segments start
    // at multiples of 256), and (-z)^256 = z^256, so the Horner pass above is shared: only the lane weights change sign.
    const int slots = paired ? 2 * ZKB_POLY_MAXPTS : ZKB_POLY_MAXPTS;
    for (int p = 0; p < pts.count; p++) {
        F zt = pts.z[p].pow_u64(t);                 // Montgomery
        F term = acc[p] * zt;
        F s = block_sum<F>(term, sh);
        F s2 = F::zero();
        if (paired) s2 = block_sum<F>((t & 1) ? term.neg() : term, sh);
        if (t == 0) {
            if (scale_by_start) {
                F zs = pts.z[p].pow_u64(start);
                s = s * zs;
                if (paired) s2 = s2 * zs;           // start is even: (-z)^start = z^start
            }
            partial[((uint64_t)b * segs + seg) * slots + p] = s;
            if (paired) partial[((uint64_t)b * segs + seg) * slots + ZKB_POLY_MAXPTS + p] = s2;
        }
    }
}

template <class P>
__global__ void poly_eval_reduce_kernel(const Fp<P> *partial, uint32_t segs, uint32_t batch, int count, uint32_t npoints_total,
                                        uint32_t point0, int paired, Fp<P> *out) {
    typedef Fp<P> F;
    const int per = paired ? 2 : 1, slots = per * ZKB_POLY_MAXPTS;
    uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= batch * (uint32_t)count * per) return;
    uint32_t b = idx / (count * per), r = idx % (count * per), p = r / per, neg = r % per;
    F s = F::zero();
    for (uint32_t g = 0; g < segs; g++) s = s + partial[((uint64_t)b * segs + g) * slots + neg * ZKB_POLY_MAXPTS + p];
    out[((uint64_t)b * npoints_total + point0 + p) * per + neg] = s;   // paired: [b][point][z, -z]
}

static void pick_segments(uint64_t n, uint32_t batch, int sm_count, uint64_t *seg_len, uint32_t *segs) {
    // enough blocks to fill the machine a few times over, segments a multiple of the block size
    uint64_t want = (uint64_t)(sm_count > 0 ? sm_count : 148) * 8;
    uint64_t s = (want + batch - 1) / batch;
    if (s < 1) s = 1;
    uint64_t len = (n + s - 1) / s;
    len = (len + ZKB_POLY_THREADS - 1) / ZKB_POLY_THREADS * ZKB_POLY_THREADS;
    if (len < 4 * ZKB_POLY_THREADS) len = 4 * ZKB_POLY_THREADS;
    *seg_len = len;
    *segs = (uint32_t)((n + len - 1) / len);
}

template <class P>
static int poly_evaluate_t(zkb_ctx *ctx, uint64_t n, uint32_t batch, const void *d_coeffs, uint32_t npoints,
                           const uint32_t *points, uint32_t *out, int paired, cudaStream_t st) {
    typedef Fp<P> F;
    const int per = paired ? 2 : 1;
    uint64_t seg_len;
    uint32_t segs;
    pick_segments(n, batch, ctx->sm_count, &seg_len, &segs);
    void *partial, *dout;
    ZKB_TRY(ctx_scratch(ctx, "eval_partial", (size_t)batch * segs * per * ZKB_POLY_MAXPTS * sizeof(F), &partial));
    ZKB_TRY(ctx_scratch(ctx, "eval_out", (size_t)batch * npoints * per * sizeof(F), &dout));
    for (uint32_t p0 = 0; p0 < npoints; p0 += ZKB_POLY_MAXPTS) {
        EvalPoints<F> pts;
        pts.count = npoints - p0 < ZKB_POLY_MAXPTS ? (int)(npoints - p0) : ZKB_POLY_MAXPTS;
        for (int p = 0; p < ZKB_POLY_MAXPTS; p++) {
            if (p < pts.count) {
                const uint32_t *src = points + (size_t)(p0 + p) * F::N;
                if (!canonical<F, P>(src)) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "evaluation point >= modulus");
                F z;
                memcpy(z.l, src, sizeof(z.l));
                pts.z[p] = z.to_mont();
                pts.z256[p] = pts.z[p].pow_u64(ZKB_POLY_THREADS);
            } else {
                pts.z[p] = F::zero();
                pts.z256[p] = F::zero();
            }
        }
        dim3 grid(segs, batch);
        poly_eval_segments_kernel<P><<<grid, ZKB_POLY_THREADS, 0, st>>>((const F *)d_coeffs, n, n, seg_len, segs, pts, 1, paired, (F *)partial);
        uint32_t total = batch * (uint32_t)pts.count * per;
        poly_eval_reduce_kernel<P><<<(total + 127) / 128, 128, 0, st>>>((const F *)partial, segs, batch, pts.count, npoints, p0, paired, (F *)dout);
        ctx->launches += 2;
        ZKB_CUDA_OK(ctx, cudaGetLastError());
    }
    ZKB_CUDA_OK(ctx, cudaMemcpyAsync(out, dout, (size_t)batch * npoints * per * sizeof(F), cudaMemcpyDeviceToHost, st));
    ZKB_CUDA_OK(ctx, cudaStreamSynchronize(st));
    return ZKB_OK;
}

// ------------------------------------------------------------------------------------ linear combination
// out[i] (+)= sum_j s_j polys[idx_j][i]  -  [i == 0] constant      (s_j Montgomery, everything else canonical)
template <class P>
__global__ void __launch_bounds__(256) poly_lincomb_kernel(const Fp<P> *polys, uint64_t n, uint64_t poly_stride, const Fp<P> *scalars,
                                                           const uint32_t *index, uint32_t terms, Fp<P> constant, int accumulate,
                                                           Fp<P> *out) {
    typedef Fp<P> F;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F acc = accumulate ? out[i] : F::zero();
    for (uint32_t j = 0; j < terms; j++) acc = acc + polys[(uint64_t)index[j] * poly_stride + i] * scalars[j];
    if (i == 0) acc = acc - constant;
    out[i] = acc;
}

template <class P>
static int poly_lincomb_t(zkb_ctx *ctx, uint64_t n, uint32_t batch, const void *d_polys, const uint32_t *scalars,
                          const uint32_t *constant, void *d_out, int accumulate, cudaStream_t st) {
    typedef Fp<P> F;
    std::vector<F> sc;
    std::vector<uint32_t> idx;
    for (uint32_t j = 0; j < batch; j++) {
        const uint32_t *src = scalars + (size_t)j * F::N;
        if (!canonical<F, P>(src)) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "lincomb scalar >= modulus");
        F s;
        memcpy(s.l, src, sizeof(s.l));
        if (s.is_zero()) continue;   // polynomial not opened at this point (lpc.hpp:143-144)
        sc.push_back(s.to_mont());
        idx.push_back(j);
    }
    F c = F::zero();
    if (constant) {
        if (!canonical<F, P>(constant)) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "lincomb constant >= modulus");
        memcpy(c.l, constant, sizeof(c.l));
    }
    void *dsc = nullptr, *didx = nullptr;
    size_t terms = sc.size();
    ZKB_TRY(ctx_scratch(ctx, "lincomb_sc", (terms + 1) * sizeof(F), &dsc));
    ZKB_TRY(ctx_scratch(ctx, "lincomb_idx", (terms + 1) * sizeof(uint32_t), &didx));
    if (terms) {
        // pageable sources: the copies are staged before the call returns, so the vectors may die afterwards
        ZKB_CUDA_OK(ctx, cudaMemcpyAsync(dsc, sc.data(), terms * sizeof(F), cudaMemcpyHostToDevice, st));
        ZKB_CUDA_OK(ctx, cudaMemcpyAsync(didx, idx.data(), terms * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        ZKB_CUDA_OK(ctx, cudaStreamSynchronize(st));
    }
    uint64_t blocks = (n + 255) / 256;
    poly_lincomb_kernel<P><<<(unsigned)blocks, 256, 0, st>>>((const F *)d_polys, n, n, (const F *)dsc, (const uint32_t *)didx,
                                                             (uint32_t)terms, c, accumulate, (F *)d_out);
    ctx->launches++;
    ZKB_CUDA_OK(ctx, cudaGetLastError());
    return ZKB_OK;
}

// ------------------------------------------------------------------------------------ division by (X - z)
// q_{i-1} = T_i with T_i = c_i + z T_{i+1}, T_n = 0 (synthetic division; T_0 = c(z) is the remainder).
// A block owns DIV_BLOCK consecutive coefficients, a thread DIV_CHUNK of them:
//   pass 1 (block sums)   S_b = sum_{k in block} c_k z^(k - block_start)
//   pass 2 (one thread)   carry_b = T_{block_end(b)} = S_{b+1} + z^DIV_BLOCK carry_{b+1}
//   pass 3                in-block suffix scan of the chunk sums (Hillis-Steele with z^(CHUNK 2^s)), then each thread
//                         runs the recurrence down its chunk and writes q.
#define ZKB_DIV_CHUNK 8
#define ZKB_DIV_BLOCK (ZKB_DIV_CHUNK * ZKB_POLY_THREADS)

template <class F>
__device__ F div_chunk_sum(const F *c, uint64_t lo, uint64_t hi, const F &z) {   // sum_{k in [lo,hi)} c_k z^(k-lo)
    F h = F::zero();
    for (uint64_t k = hi; k-- > lo;) h = h * z + c[k];
    return h;
}

// V_t = sum_{t' >= t} H_t' zc^(t'-t) over the threads of the block (zc = z^CHUNK, Montgomery)
template <class F>
__device__ F div_block_suffix(F h, F zc, F *sh) {
    const uint32_t t = threadIdx.x;
    F v = h, m = zc;
    for (uint32_t s = 1; s < ZKB_POLY_THREADS; s <<= 1) {
        sh[t] = v;
        __syncthreads();
        if (t + s < ZKB_POLY_THREADS) v = v + sh[t + s] * m;
        __syncthreads();
        m = m * m;
    }
    return v;
}

template <class P>
__global__ void __launch_bounds__(ZKB_POLY_THREADS) poly_div_sums_kernel(const Fp<P> *c, uint64_t n, Fp<P> z, Fp<P> zc, Fp<P> *sums) {
    typedef Fp<P> F;
    __shared__ F sh[ZKB_POLY_THREADS];
    const uint64_t start = (uint64_t)blockIdx.x * ZKB_DIV_BLOCK;
    uint64_t lo = start + (uint64_t)threadIdx.x * ZKB_DIV_CHUNK, hi = lo + ZKB_DIV_CHUNK;
    if (lo > n) lo = n;
    if (hi > n) hi = n;
    F v = div_block_suffix<F>(div_chunk_sum<F>(c, lo, hi, z), zc, sh);
    if (threadIdx.x == 0) sums[blockIdx.x] = v;
}

template <class P>
__global__ void poly_div_carry_kernel(const Fp<P> *sums, uint32_t blocks, Fp<P> zb, Fp<P> *carry, Fp<P> *remainder) {
    typedef Fp<P> F;
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    F acc = F::zero();                       // T at the end of the last block
    for (uint32_t b = blocks; b-- > 0;) {
        carry[b] = acc;
        acc = sums[b] + acc * zb;
    }
    *remainder = acc;                        // T_0 = c(z)
}

template <class P>
__global__ void __launch_bounds__(ZKB_POLY_THREADS) poly_div_apply_kernel(const Fp<P> *c, uint64_t n, Fp<P> z, Fp<P> zc,
                                                                          const Fp<P> *carry, Fp<P> *q) {
    typedef Fp<P> F;
    __shared__ F sh[ZKB_POLY_THREADS + 1];
    const uint32_t t = threadIdx.x;
    const uint64_t start = (uint64_t)blockIdx.x * ZKB_DIV_BLOCK;
    uint64_t lo = start + (uint64_t)t * ZKB_DIV_CHUNK, hi = lo + ZKB_DIV_CHUNK;
    if (lo > n) lo = n;
    if (hi > n) hi = n;
    F v = div_block_suffix<F>(div_chunk_sum<F>(c, lo, hi, z), zc, sh);
    // T at the end of my chunk = V_{t+1} + carry_b zc^(255 - t)   (V_256 = 0)
    __syncthreads();
    sh[t] = v;
    if (t == 0) sh[ZKB_POLY_THREADS] = F::zero();
    __syncthreads();
    F T = sh[t + 1] + carry[blockIdx.x] * zc.pow_u64(ZKB_POLY_THREADS - 1 - t);
    for (uint64_t k = hi; k-- > lo;) {
        T = T * z + c[k];                    // T_k
        if (k > 0) q[k - 1] = T;
    }
    if (hi == n && lo < n) q[n - 1] = F::zero();
}

template <class P>
static int poly_div_linear_t(zkb_ctx *ctx, uint64_t n, const void *d_in, const uint32_t *point, void *d_out,
                             uint32_t *remainder_out, cudaStream_t st) {
    typedef Fp<P> F;
    if (!canonical<F, P>(point)) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "division point >= modulus");
    F z;
    memcpy(z.l, point, sizeof(z.l));
    z = z.to_mont();
    F zc = z.pow_u64(ZKB_DIV_CHUNK), zb = z.pow_u64(ZKB_DIV_BLOCK);
    uint32_t blocks = (uint32_t)((n + ZKB_DIV_BLOCK - 1) / ZKB_DIV_BLOCK);
    void *sums;
    ZKB_TRY(ctx_scratch(ctx, "div_sums", (size_t)(2 * blocks + 1) * sizeof(F), &sums));
    F *d_sums = (F *)sums, *d_carry = d_sums + blocks, *d_rem = d_carry + blocks;
    poly_div_sums_kernel<P><<<blocks, ZKB_POLY_THREADS, 0, st>>>((const F *)d_in, n, z, zc, d_sums);
    poly_div_carry_kernel<P><<<1, 32, 0, st>>>(d_sums, blocks, zb, d_carry, d_rem);
    poly_div_apply_kernel<P><<<blocks, ZKB_POLY_THREADS, 0, st>>>((const F *)d_in, n, z, zc, d_carry, (F *)d_out);
    ctx->launches += 3;
    ZKB_CUDA_OK(ctx, cudaGetLastError());
    if (remainder_out) {
        ZKB_CUDA_OK(ctx, cudaMemcpyAsync(remainder_out, d_rem, sizeof(F), cudaMemcpyDeviceToHost, st));
        ZKB_CUDA_OK(ctx, cudaStreamSynchronize(st));
    }
    return ZKB_OK;
}

// ------------------------------------------------------------------------------------ sparse matrix x vector
// Rows longer than ZKB_SPMV_LONG terms (the closing constraint of r1cs_examples.hpp:121-129 sums every variable) are
// cut into segments of ZKB_SPMV_SEG terms, one block each, and the segment sums are added per row afterwards.
#define ZKB_SPMV_LONG 512
#define ZKB_SPMV_SEG 4096
struct zkb_sparse_matrix {
    zkb_ctx *ctx;
    int device;   // freed with plain cudaFree on this device: the matrix may outlive its context
    int field;
    uint64_t rows, cols, nnz;
    uint64_t *d_row_ptr;
    uint32_t *d_col;
    void *d_val;   // Montgomery form
    uint32_t n_long, n_segs;
    uint64_t *d_seg;        // [n_segs][2]: begin, end (positions in col/val)
    uint64_t *d_long;       // [n_long][3]: row, first segment, segment count
    void *d_partial;        // [n_segs] elements
};

template <class P>
__global__ void __launch_bounds__(256) to_mont_kernel(uint64_t n, Fp<P> *v) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) v[i] = v[i].to_mont();
}

// one thread per row: R1CS rows are short linear combinations (r1cs_examples.hpp:77-146 has <= 3 terms per row)
template <class P>
__global__ void __launch_bounds__(256) spmv_kernel(uint64_t rows, const uint64_t *row_ptr, const uint32_t *col, const Fp<P> *val,
                                                   const Fp<P> *x, Fp<P> *y) {
    typedef Fp<P> F;
    uint64_t r = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= rows) return;
    const uint64_t k0 = row_ptr[r], k1 = row_ptr[r + 1];
    if (k1 - k0 > ZKB_SPMV_LONG) return;   // handled by the segment kernels
    F acc = F::zero();
    for (uint64_t k = k0; k < k1; k++) acc = acc + x[col[k]] * val[k];
    y[r] = acc;
}

template <class P>
__global__ void __launch_bounds__(ZKB_POLY_THREADS) spmv_segment_kernel(const uint64_t *seg, const uint32_t *col, const Fp<P> *val,
                                                                        const Fp<P> *x, Fp<P> *partial) {
    typedef Fp<P> F;
    __shared__ F sh[ZKB_POLY_THREADS];
    const uint64_t k0 = seg[2 * blockIdx.x], k1 = seg[2 * blockIdx.x + 1];
    F acc = F::zero();
    for (uint64_t k = k0 + threadIdx.x; k < k1; k += ZKB_POLY_THREADS) acc = acc + x[col[k]] * val[k];
    F s = block_sum<F>(acc, sh);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

template <class P>
__global__ void spmv_long_rows_kernel(uint32_t n_long, const uint64_t *lrows, const Fp<P> *partial, Fp<P> *y) {
    typedef Fp<P> F;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_long) return;
    F acc = F::zero();
    for (uint64_t g = 0; g < lrows[3 * i + 2]; g++) acc = acc + partial[lrows[3 * i + 1] + g];
    y[lrows[3 * i]] = acc;
}

template <class P>
static int sparse_to_mont_t(zkb_ctx *ctx, uint64_t nnz, void *d_val, cudaStream_t st) {
    if (nnz == 0) return ZKB_OK;
    to_mont_kernel<P><<<(unsigned)((nnz + 255) / 256), 256, 0, st>>>(nnz, (Fp<P> *)d_val);
    ctx->launches++;
    ZKB_CUDA_OK(ctx, cudaGetLastError());
    return ZKB_OK;
}
template <class P>
static int spmv_t(zkb_ctx *ctx, const zkb_sparse_matrix *m, const void *dx, void *dy, cudaStream_t st) {
    if (m->rows == 0) return ZKB_OK;
    spmv_kernel<P><<<(unsigned)((m->rows + 255) / 256), 256, 0, st>>>(m->rows, m->d_row_ptr, m->d_col, (const Fp<P> *)m->d_val,
                                                                     (const Fp<P> *)dx, (Fp<P> *)dy);
    ctx->launches++;
    if (m->n_long) {
        spmv_segment_kernel<P><<<m->n_segs, ZKB_POLY_THREADS, 0, st>>>(m->d_seg, m->d_col, (const Fp<P> *)m->d_val, (const Fp<P> *)dx,
                                                                       (Fp<P> *)m->d_partial);
        spmv_long_rows_kernel<P><<<(m->n_long + 127) / 128, 128, 0, st>>>(m->n_long, m->d_long, (const Fp<P> *)m->d_partial, (Fp<P> *)dy);
        ctx->launches += 2;
    }
    ZKB_CUDA_OK(ctx, cudaGetLastError());
    return ZKB_OK;
}

static int poly_evaluate_dispatch(zkb_ctx *ctx, int field, uint64_t n, uint32_t batch, const void *d, uint32_t npoints,
                                  const uint32_t *points, uint32_t *out, int paired, cudaStream_t st) {
    ZKB_DISPATCH_FR(field, poly_evaluate_t, ctx, n, batch, d, npoints, points, out, paired, st)
}
static int poly_lincomb_dispatch(zkb_ctx *ctx, int field, uint64_t n, uint32_t batch, const void *d, const uint32_t *sc,
                                 const uint32_t *c, void *out, int acc, cudaStream_t st) {
    ZKB_DISPATCH_FR(field, poly_lincomb_t, ctx, n, batch, d, sc, c, out, acc, st)
}
static int poly_div_dispatch(zkb_ctx *ctx, int field, uint64_t n, const void *d, const uint32_t *pt, void *out, uint32_t *rem,
                             cudaStream_t st) {
    ZKB_DISPATCH_FR(field, poly_div_linear_t, ctx, n, d, pt, out, rem, st)
}
static int sparse_to_mont_dispatch(zkb_ctx *ctx, int field, uint64_t nnz, void *d, cudaStream_t st) {
    ZKB_DISPATCH_FR(field, sparse_to_mont_t, ctx, nnz, d, st)
}
static int spmv_dispatch(zkb_ctx *ctx, int field, const zkb_sparse_matrix *m, const void *dx, void *dy, cudaStream_t st) {
    ZKB_DISPATCH_FR(field, spmv_t, ctx, m, dx, dy, st)
}

extern "C" {

int zkb_poly_evaluate(zkb_ctx *ctx, int field, int form, uint64_t n, uint32_t batch, const void *polys, int mem,
                      uint32_t npoints, const uint32_t *points, uint32_t *out, void *stream) {
    if (!ctx) return ZKB_ERR_INVALID_ARGUMENT;
    if (!is_fr(field) || n == 0 || (form != ZKB_POLY_COEFFICIENTS && form != ZKB_POLY_DFS) || !polys || !points || !out)
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_poly_evaluate: bad arguments");
    if (batch == 0 || npoints == 0) return ZKB_OK;
    int log_n = 0;
    if (form == ZKB_POLY_DFS) {
        while ((1ull << log_n) < n) log_n++;
        if ((1ull << log_n) != n) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_poly_evaluate: dfs size must be a power of two");
        if (log_n > zkb_field_two_adicity(field)) return ctx_fail(ctx, ZKB_ERR_DOMAIN_TOO_LARGE, "2^log_n exceeds the two-adicity");
    }
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t bytes = (size_t)batch * n * 32;
    const void *d = polys;
    if (mem != ZKB_MEM_DEVICE) {
        void *p;
        ZKB_TRY(ctx_scratch(ctx, "io_in", bytes, &p));
        ZKB_CUDA_OK(ctx, cudaMemcpyAsync(p, polys, bytes, cudaMemcpyHostToDevice, st));
        d = p;
    }
    if (form == ZKB_POLY_DFS) {   // polynomial_dfs::evaluate = coefficients(), then the value
        void *coef;
        ZKB_TRY(ctx_scratch(ctx, "eval_coef", bytes, &coef));
        ZKB_TRY(ntt_device(ctx, field, log_n, batch, d, coef, 1, nullptr, n, n, st));
        d = coef;
    }
    return poly_evaluate_dispatch(ctx, field, n, batch, d, npoints, points, out, 0, st);
}

int zkb_poly_evaluate_pm(zkb_ctx *ctx, int field, uint64_t n, uint32_t batch, const void *polys_device, uint32_t npoints,
                         const uint32_t *points, uint32_t *out, void *stream) {
    if (!ctx) return ZKB_ERR_INVALID_ARGUMENT;
    if (!is_fr(field) || n == 0 || !polys_device || !points || !out)
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_poly_evaluate_pm: bad arguments");
    if (batch == 0 || npoints == 0) return ZKB_OK;
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    return poly_evaluate_dispatch(ctx, field, n, batch, polys_device, npoints, points, out, 1, (cudaStream_t)stream);
}

int zkb_poly_lincomb(zkb_ctx *ctx, int field, uint64_t n, uint32_t batch, const void *polys_device, const uint32_t *scalars,
                     const uint32_t *constant, void *out_device, int accumulate, void *stream) {
    if (!ctx) return ZKB_ERR_INVALID_ARGUMENT;
    if (!is_fr(field) || n == 0 || !out_device || (batch && (!polys_device || !scalars)))
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_poly_lincomb: bad arguments");
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    return poly_lincomb_dispatch(ctx, field, n, batch, polys_device, scalars, constant, out_device, accumulate, (cudaStream_t)stream);
}

int zkb_poly_div_linear(zkb_ctx *ctx, int field, uint64_t n, const void *in_device, const uint32_t *point, void *out_device,
                        uint32_t *remainder_out, void *stream) {
    if (!ctx) return ZKB_ERR_INVALID_ARGUMENT;
    if (!is_fr(field) || n == 0 || !in_device || !out_device || !point || in_device == out_device)
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_poly_div_linear: bad arguments (in place is not supported)");
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    return poly_div_dispatch(ctx, field, n, in_device, point, out_device, remainder_out, (cudaStream_t)stream);
}

int zkb_sparse_matrix_create(zkb_ctx *ctx, int field, uint64_t rows, uint64_t cols, const uint64_t *row_ptr,
                             const uint32_t *col_idx, const uint32_t *values, void *stream, zkb_sparse_matrix **out) {
    if (!ctx || !out) return ZKB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (!is_fr(field) || !row_ptr || cols == 0 || cols > 0xffffffffull)
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_sparse_matrix_create: bad arguments");
    const uint64_t nnz = row_ptr[rows];
    if (row_ptr[0] != 0 || (nnz && (!col_idx || !values)))
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_sparse_matrix_create: bad CSR arrays");
    for (uint64_t r = 0; r < rows; r++)
        if (row_ptr[r + 1] < row_ptr[r]) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_sparse_matrix_create: row_ptr not monotone");
    for (uint64_t k = 0; k < nnz; k++)
        if (col_idx[k] >= cols) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_sparse_matrix_create: column index out of range");
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    zkb_sparse_matrix *m = new zkb_sparse_matrix();
    m->ctx = ctx; m->device = ctx->device; m->field = field; m->rows = rows; m->cols = cols; m->nnz = nnz;
    m->d_row_ptr = nullptr; m->d_col = nullptr; m->d_val = nullptr;
    m->d_seg = nullptr; m->d_long = nullptr; m->d_partial = nullptr;
    std::vector<uint64_t> segs, longs;
    for (uint64_t r = 0; r < rows; r++) {
        const uint64_t k0 = row_ptr[r], k1 = row_ptr[r + 1];
        if (k1 - k0 <= ZKB_SPMV_LONG) continue;
        longs.push_back(r);
        longs.push_back(segs.size() / 2);
        longs.push_back((k1 - k0 + ZKB_SPMV_SEG - 1) / ZKB_SPMV_SEG);
        for (uint64_t k = k0; k < k1; k += ZKB_SPMV_SEG) {
            segs.push_back(k);
            segs.push_back(k + ZKB_SPMV_SEG < k1 ? k + ZKB_SPMV_SEG : k1);
        }
    }
    m->n_long = (uint32_t)(longs.size() / 3);
    m->n_segs = (uint32_t)(segs.size() / 2);
    cudaError_t e = cudaMalloc((void **)&m->d_row_ptr, (rows + 1) * sizeof(uint64_t));
    if (e == cudaSuccess && m->n_long) {
        e = cudaMalloc((void **)&m->d_seg, segs.size() * sizeof(uint64_t));
        if (e == cudaSuccess) e = cudaMalloc((void **)&m->d_long, longs.size() * sizeof(uint64_t));
        if (e == cudaSuccess) e = cudaMalloc(&m->d_partial, (size_t)m->n_segs * 32);
        if (e == cudaSuccess) e = cudaMemcpyAsync(m->d_seg, segs.data(), segs.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, st);
        if (e == cudaSuccess) e = cudaMemcpyAsync(m->d_long, longs.data(), longs.size() * sizeof(uint64_t), cudaMemcpyHostToDevice, st);
    }
    if (e == cudaSuccess) e = cudaMalloc((void **)&m->d_col, (nnz + 1) * sizeof(uint32_t));
    if (e == cudaSuccess) e = cudaMalloc(&m->d_val, (nnz + 1) * 32);
    if (e == cudaSuccess) e = cudaMemcpyAsync(m->d_row_ptr, row_ptr, (rows + 1) * sizeof(uint64_t), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(m->d_col, col_idx, nnz * sizeof(uint32_t), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && nnz) e = cudaMemcpyAsync(m->d_val, values, nnz * 32, cudaMemcpyHostToDevice, st);
    int s = ZKB_OK;
    if (e != cudaSuccess) {
        cudaGetLastError();
        s = ctx_fail(ctx, e == cudaErrorMemoryAllocation ? ZKB_ERR_OUT_OF_MEMORY : ZKB_ERR_CUDA,
                     std::string("zkb_sparse_matrix_create: ") + cudaGetErrorString(e));
    }
    if (s == ZKB_OK) s = sparse_to_mont_dispatch(ctx, field, nnz, m->d_val, st);
    if (s == ZKB_OK && cudaStreamSynchronize(st) != cudaSuccess) s = ctx_fail(ctx, ZKB_ERR_CUDA, "zkb_sparse_matrix_create: sync failed");
    if (s != ZKB_OK) {
        zkb_sparse_matrix_free(m);
        return s;
    }
    *out = m;
    return ZKB_OK;
}

void zkb_sparse_matrix_free(zkb_sparse_matrix *m) {
    if (!m) return;
    cudaSetDevice(m->device);
    if (m->d_row_ptr) cudaFree(m->d_row_ptr);
    if (m->d_col) cudaFree(m->d_col);
    if (m->d_val) cudaFree(m->d_val);
    if (m->d_seg) cudaFree(m->d_seg);
    if (m->d_long) cudaFree(m->d_long);
    if (m->d_partial) cudaFree(m->d_partial);
    delete m;
}

int zkb_sparse_matvec(zkb_ctx *ctx, const zkb_sparse_matrix *m, const void *x, int x_mem, void *y_device, void *stream) {
    if (!ctx || !m) return ZKB_ERR_INVALID_ARGUMENT;
    if (!x || !y_device) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_sparse_matvec: null vector");
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const void *dx = x;
    if (x_mem != ZKB_MEM_DEVICE) {
        void *p;
        ZKB_TRY(ctx_scratch(ctx, "spmv_x", m->cols * 32, &p));
        ZKB_CUDA_OK(ctx, cudaMemcpyAsync(p, x, m->cols * 32, cudaMemcpyHostToDevice, st));
        dx = p;
    }
    return spmv_dispatch(ctx, m->field, m, dx, y_device, st);
}

// ------------------------------------------------------------------------------------ FRI commit phase
int zkb_fri_commit_phase(zkb_ctx *ctx, int field, int hash, int log_n, const void *f, int mem, const uint32_t *step_list,
                         uint32_t rounds, zkb_fri_challenge_fn challenge, void *user, uint8_t *roots_out,
                         zkb_merkle_tree **trees_out, void *fs_device_out, uint32_t *alphas_out, uint32_t *final_poly_out,
                         void *stream) {
    if (!ctx) return ZKB_ERR_INVALID_ARGUMENT;
    const int db = zkb_merkle_digest_bytes(hash);
    if (trees_out)
        for (uint32_t i = 0; i < rounds; i++) trees_out[i] = nullptr;
    if (!is_fr(field) || !db || !f || !step_list || rounds == 0 || !challenge || !roots_out || log_n < 1)
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_fri_commit_phase: bad arguments");
    if (log_n > zkb_field_two_adicity(field)) return ctx_fail(ctx, ZKB_ERR_DOMAIN_TOO_LARGE, "2^log_n exceeds the two-adicity");
    uint32_t total = 0, max_step = 0;
    for (uint32_t i = 0; i < rounds; i++) {
        if (step_list[i] < 1) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_fri_commit_phase: step_list entries must be >= 1");
        if ((int)(total + step_list[i]) > log_n)
            return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_fri_commit_phase: sum(step_list) exceeds log2 |D0|");
        total += step_list[i];
        if (step_list[i] > max_step) max_step = step_list[i];
    }
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const size_t fbytes = ((size_t)32) << log_n;
    const void *cur = f;
    if (mem != ZKB_MEM_DEVICE) {
        void *p;
        ZKB_TRY(ctx_scratch(ctx, "io_in", fbytes, &p));
        ZKB_CUDA_OK(ctx, cudaMemcpyAsync(p, f, fbytes, cudaMemcpyHostToDevice, st));
        cur = p;
    }
    void *pp[2];
    ZKB_TRY(ctx_scratch(ctx, "fri_a", fbytes / 2, &pp[0]));
    ZKB_TRY(ctx_scratch(ctx, "fri_b", fbytes / 2, &pp[1]));
    int which = 0, cur_log = log_n, status = ZKB_OK;
    char *fs = (char *)fs_device_out;
    std::vector<uint32_t> alphas((size_t)max_step * 8);
    uint32_t t = 0;
    for (uint32_t i = 0; i < rounds && status == ZKB_OK; i++) {
        // precommit<FRI>(f, D[t], step_list[i]) -> root -> transcript (basic_fri.hpp:715-719, 731-732)
        uint8_t *root = roots_out + (size_t)i * db;
        status = merkle_build_device(ctx, hash, cur_log, (int)step_list[i], 1, cur, root, trees_out ? &trees_out[i] : nullptr, st);
        if (status != ZKB_OK) break;
        if (challenge(user, i, root, (uint32_t)db, step_list[i], alphas.data()) != 0) {
            status = ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_fri_commit_phase: the challenge callback failed");
            break;
        }
        for (uint32_t s = 0; s < step_list[i] && status == ZKB_OK; s++, t++) {
            const uint32_t *alpha = alphas.data() + (size_t)s * 8;
            if (alphas_out) memcpy(alphas_out + (size_t)t * 8, alpha, 32);
            void *dst;
            if (fs && s + 1 == step_list[i]) {
                dst = fs;
                fs += ((size_t)32) << (cur_log - 1);
            } else {
                dst = pp[which];
                which ^= 1;
            }
            status = fold_device(ctx, field, cur_log, cur, alpha, dst, st);   // fold_polynomial(f, alphas[t], D[t])
            cur = dst;
            cur_log--;
        }
    }
    if (status == ZKB_OK && final_poly_out) {
        // final_polynomial = f.coefficients() (basic_fri.hpp:735-737)
        void *coef = pp[which];
        status = ntt_device(ctx, field, cur_log, 1, cur, coef, 1, nullptr, 1ull << cur_log, 1ull << cur_log, st);
        if (status == ZKB_OK) {
            cudaError_t e = cudaMemcpyAsync(final_poly_out, coef, ((size_t)32) << cur_log, cudaMemcpyDeviceToHost, st);
            if (e != cudaSuccess) status = ctx_fail(ctx, ZKB_ERR_CUDA, cudaGetErrorString(e));
        }
    }
    cudaError_t e = cudaStreamSynchronize(st);
    if (status == ZKB_OK && e != cudaSuccess) status = ctx_fail(ctx, ZKB_ERR_CUDA, cudaGetErrorString(e));
    if (status != ZKB_OK && trees_out)
        for (uint32_t i = 0; i < rounds; i++) {
            zkb_merkle_free(trees_out[i]);
            trees_out[i] = nullptr;
        }
    return status;
}

}  // extern "C"
