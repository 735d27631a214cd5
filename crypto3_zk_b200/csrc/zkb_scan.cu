// Multiplicative scans over field elements: prefix products, batch inversion and the grand product V_P of the
// Placeholder permutation argument.
//
// Reference (the first piece of the argument builders):
//   zk/snark/systems/plonk/placeholder/permutation_argument.hpp:104-133
//     g_v[i] = column_i + beta S_id[i] + gamma,  h_v[i] = column_i + beta S_sigma[i] + gamma,
//     V_P[0] = 1,  V_P[j] = V_P[j-1] * prod_i g_v[i][j-1] * (prod_i h_v[i][j-1]).inversed()
//   - one field inversion per row on the CPU.  Here: the row products in one pass, ONE inversion for the whole
//   column (Montgomery's trick as two scans: 1/h_j = P_j S_j / T with P / S the exclusive prefix / suffix products and
//   T the total), one exclusive prefix-product scan for V_P.
//
// Scan = three launches: block totals (each thread multiplies a chunk of 8, the block multiplies its 256 chunk
// products), one block that scans the block totals, and an apply pass that redoes the chunk products with the carries.
// Bound: integer pipe (3 products per element and scan).  Data are canonical at the ABI, Montgomery inside.
#include <string.h>
#include "zkb_field.cuh"
#include "zkb_internal.h"

using namespace zkb;

#define SCAN_THREADS 256
#define SCAN_CHUNK 8
#define SCAN_BLOCK (SCAN_THREADS * SCAN_CHUNK)

// element i of the scan order
__device__ __forceinline__ uint64_t scan_idx(uint64_t i, uint64_t n, int reverse) { return reverse ? n - 1 - i : i; }

// exclusive scan (by products) of one value per thread; *total receives the product of all 256.  All threads call.
template <class F>
__device__ F block_exclusive_prod(F v, F *sh /* 2 * SCAN_THREADS */, F *total) {
    const uint32_t t = threadIdx.x;
    F *a = sh, *b = sh + SCAN_THREADS;
    a[t] = v;
    __syncthreads();
    for (uint32_t d = 1; d < SCAN_THREADS; d <<= 1) {      // Hillis-Steele, inclusive
        F x = a[t];
        if (t >= d) x = a[t - d] * x;
        b[t] = x;
        __syncthreads();
        F *s = a; a = b; b = s;
    }
    *total = a[SCAN_THREADS - 1];
    F r = t == 0 ? F::one() : a[t - 1];
    __syncthreads();
    return r;
}

template <class P>
__global__ void __launch_bounds__(SCAN_THREADS) scan_totals_kernel(const Fp<P> *__restrict__ in, uint64_t n, int reverse,
                                                                   Fp<P> *__restrict__ totals) {
    typedef Fp<P> F;
    __shared__ F sh[2 * SCAN_THREADS];
    const uint64_t i0 = (uint64_t)blockIdx.x * SCAN_BLOCK + (uint64_t)threadIdx.x * SCAN_CHUNK;
    F acc = F::one();
#pragma unroll
    for (int k = 0; k < SCAN_CHUNK; k++)
        if (i0 + k < n) acc = acc * in[scan_idx(i0 + k, n, reverse)];
    F total;
    block_exclusive_prod<F>(acc, sh, &total);
    if (threadIdx.x == 0) totals[blockIdx.x] = total;
}

// one block: carry[b] = product of totals[0 .. b), *total = product of all; optionally its inverse (Fermat, one thread)
template <class P>
__global__ void __launch_bounds__(SCAN_THREADS) scan_carry_kernel(const Fp<P> *__restrict__ totals, uint32_t blocks,
                                                                  Fp<P> *__restrict__ carry, Fp<P> *__restrict__ total_out,
                                                                  Fp<P> *__restrict__ total_inv_out) {
    typedef Fp<P> F;
    __shared__ F sh[2 * SCAN_THREADS];
    const uint32_t per = (blocks + SCAN_THREADS - 1) / SCAN_THREADS;
    const uint32_t b0 = threadIdx.x * per;
    F acc = F::one();
    for (uint32_t k = 0; k < per; k++)
        if (b0 + k < blocks) acc = acc * totals[b0 + k];
    F total;
    F pre = block_exclusive_prod<F>(acc, sh, &total);
    for (uint32_t k = 0; k < per; k++)
        if (b0 + k < blocks) {
            carry[b0 + k] = pre;
            pre = pre * totals[b0 + k];
        }
    if (threadIdx.x == 0) {
        if (total_out) *total_out = total;
        if (total_inv_out) *total_inv_out = total.inverse();
    }
}

// out[idx(i)] = product of in[idx(j)] for j < i (exclusive) or j <= i; from_mont: store canonical
template <class P>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply_kernel(const Fp<P> *__restrict__ in, uint64_t n, int reverse, int exclusive,
                                                                  int from_mont, const Fp<P> *__restrict__ carry,
                                                                  Fp<P> *__restrict__ out) {
    typedef Fp<P> F;
    __shared__ F sh[2 * SCAN_THREADS];
    const uint64_t i0 = (uint64_t)blockIdx.x * SCAN_BLOCK + (uint64_t)threadIdx.x * SCAN_CHUNK;
    F v[SCAN_CHUNK];
    F acc = F::one();
#pragma unroll
    for (int k = 0; k < SCAN_CHUNK; k++) {
        v[k] = i0 + k < n ? in[scan_idx(i0 + k, n, reverse)] : F::one();
        acc = acc * v[k];
    }
    F total;
    F run = block_exclusive_prod<F>(acc, sh, &total) * carry[blockIdx.x];
#pragma unroll
    for (int k = 0; k < SCAN_CHUNK; k++) {
        if (i0 + k >= n) break;
        F o = run;
        run = run * v[k];
        if (!exclusive) o = run;
        out[scan_idx(i0 + k, n, reverse)] = from_mont ? o.from_mont() : o;
    }
}

template <class P>
__global__ void __launch_bounds__(256) scan_to_mont_kernel(uint64_t n, const Fp<P> *__restrict__ in, Fp<P> *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = in[i].to_mont();
}

// inv[i] = pre[i] * suf[i] * total_inv  (Montgomery in; canonical out when from_mont); optionally times num[i]
template <class P>
__global__ void __launch_bounds__(256) scan_inverse_finish_kernel(uint64_t n, const Fp<P> *__restrict__ pre, const Fp<P> *__restrict__ suf,
                                                                  const Fp<P> *__restrict__ total_inv, const Fp<P> *__restrict__ num,
                                                                  int from_mont, Fp<P> *__restrict__ out) {
    typedef Fp<P> F;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F r = pre[i] * suf[i] * *total_inv;
    if (num) r = r * num[i];
    out[i] = from_mont ? r.from_mont() : r;
}

// nom[j] = prod_i (col_i[j] + beta sid_i[j] + gamma), den[j] likewise with sigma (canonical in, Montgomery out)
template <class P>
__global__ void __launch_bounds__(256) perm_rows_kernel(uint64_t n, uint32_t ncols, const Fp<P> *__restrict__ cols,
                                                        const Fp<P> *__restrict__ sid, const Fp<P> *__restrict__ ssigma, Fp<P> beta,
                                                        Fp<P> gamma, Fp<P> *__restrict__ nom, Fp<P> *__restrict__ den) {
    typedef Fp<P> F;
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    F a = F::one(), b = F::one();
    for (uint32_t i = 0; i < ncols; i++) {
        // canonical x times Montgomery beta is the canonical product.  The canonical factors go into the Montgomery products
        // as they are: that scales nom and den by the same R^-ncols, which cancels in nom / den - no to_mont per factor.
        F c = cols[(uint64_t)i * n + j];
        F g = c + sid[(uint64_t)i * n + j] * beta + gamma;
        F h = c + ssigma[(uint64_t)i * n + j] * beta + gamma;
        a = a * g;
        b = b * h;
    }
    nom[j] = a;
    den[j] = b;
}

// lookup argument, compute_V_L (lookup_argument.hpp:375-409): for row j < usable
//   nom[j] = (1+beta)^n_in prod_i (gamma + input_i[j]) prod_i (part1 + value_i[j] + beta value_i[j+1])
//   den[j] = prod_i (part1 + sorted_i[j] + beta sorted_i[j+1]),   part1 = (1+beta) gamma;   rows >= usable: 1
template <class P>
__global__ void __launch_bounds__(256) lookup_rows_kernel(uint64_t n, uint64_t usable, uint32_t n_in, const Fp<P> *__restrict__ inputs,
                                                          uint32_t n_val, const Fp<P> *__restrict__ values, uint32_t n_sorted,
                                                          const Fp<P> *__restrict__ sorted, Fp<P> beta_mont, Fp<P> gamma, Fp<P> part1,
                                                          Fp<P> one_beta_pow_mont, Fp<P> *__restrict__ nom, Fp<P> *__restrict__ den) {
    typedef Fp<P> F;
    uint64_t j = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= n) return;
    if (j >= usable) {
        nom[j] = F::one();
        den[j] = F::one();
        return;
    }
    F a = one_beta_pow_mont, b = F::one();
    for (uint32_t i = 0; i < n_in; i++) a = a * (gamma + inputs[(uint64_t)i * n + j]).to_mont();
    for (uint32_t i = 0; i < n_val; i++)
        a = a * (part1 + values[(uint64_t)i * n + j] + values[(uint64_t)i * n + j + 1] * beta_mont).to_mont();
    for (uint32_t i = 0; i < n_sorted; i++)
        b = b * (part1 + sorted[(uint64_t)i * n + j] + sorted[(uint64_t)i * n + j + 1] * beta_mont).to_mont();
    nom[j] = a;
    den[j] = b;
}

template <class P>
static int scan_run(zkb_ctx *ctx, const Fp<P> *in, uint64_t n, int reverse, int exclusive, int from_mont, Fp<P> *out, Fp<P> *total_out,
                    Fp<P> *total_inv_out, cudaStream_t st) {
    typedef Fp<P> F;
    const uint32_t blocks = (uint32_t)((n + SCAN_BLOCK - 1) / SCAN_BLOCK);
    void *p;
    ZKB_TRY(ctx_scratch(ctx, reverse ? "scan_tot_r" : "scan_tot", (size_t)2 * blocks * sizeof(F), &p));
    F *totals = (F *)p, *carry = totals + blocks;
    scan_totals_kernel<P><<<blocks, SCAN_THREADS, 0, st>>>(in, n, reverse, totals);
    scan_carry_kernel<P><<<1, SCAN_THREADS, 0, st>>>(totals, blocks, carry, total_out, total_inv_out);
    scan_apply_kernel<P><<<blocks, SCAN_THREADS, 0, st>>>(in, n, reverse, exclusive, from_mont, carry, out);
    ctx->launches += 3;
    ZKB_CUDA_OK(ctx, cudaGetLastError());
    return ZKB_OK;
}

// inverse of every element of `den` (Montgomery), times num (or nullptr); out Montgomery or canonical.  *zero_flag (host)
// is set when some element is zero (the total product vanishes).
template <class P>
static int batch_inverse_mont(zkb_ctx *ctx, const Fp<P> *den, const Fp<P> *num, uint64_t n, int from_mont, Fp<P> *out, bool *zero_found,
                              cudaStream_t st) {
    typedef Fp<P> F;
    void *p;
    ZKB_TRY(ctx_scratch(ctx, "scan_inv", (size_t)(2 * n + 2) * sizeof(F), &p));
    F *pre = (F *)p, *suf = pre + n, *tot = suf + n, *tinv = tot + 1;
    ZKB_TRY(scan_run<P>(ctx, den, n, 0, 1, 0, pre, tot, tinv, st));
    ZKB_TRY(scan_run<P>(ctx, den, n, 1, 1, 0, suf, nullptr, nullptr, st));
    F h;
    ZKB_CUDA_OK(ctx, cudaMemcpyAsync(&h, tot, sizeof(F), cudaMemcpyDeviceToHost, st));
    ZKB_CUDA_OK(ctx, cudaStreamSynchronize(st));
    bool z = true;
    for (int i = 0; i < F::N; i++) z = z && h.l[i] == 0;
    *zero_found = z;
    if (z) return ZKB_OK;
    scan_inverse_finish_kernel<P><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, pre, suf, tinv, num, from_mont, out);
    ctx->launches++;
    ZKB_CUDA_OK(ctx, cudaGetLastError());
    return ZKB_OK;
}

template <class P>
static int to_mont_copy(zkb_ctx *ctx, const void *in, uint64_t n, const char *role, Fp<P> **out, cudaStream_t st) {
    void *p;
    ZKB_TRY(ctx_scratch(ctx, role, (size_t)n * sizeof(Fp<P>), &p));
    scan_to_mont_kernel<P><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, (const Fp<P> *)in, (Fp<P> *)p);
    ctx->launches++;
    ZKB_CUDA_OK(ctx, cudaGetLastError());
    *out = (Fp<P> *)p;
    return ZKB_OK;
}

template <class P>
static int prefix_product_t(zkb_ctx *ctx, uint64_t n, const void *in, void *out, int exclusive, cudaStream_t st) {
    Fp<P> *m;
    ZKB_TRY(to_mont_copy<P>(ctx, in, n, "scan_in", &m, st));
    return scan_run<P>(ctx, m, n, 0, exclusive, 1, (Fp<P> *)out, nullptr, nullptr, st);
}

template <class P>
static int batch_inverse_t(zkb_ctx *ctx, uint64_t n, const void *in, void *out, cudaStream_t st) {
    Fp<P> *m;
    ZKB_TRY(to_mont_copy<P>(ctx, in, n, "scan_in", &m, st));
    bool zero = false;
    ZKB_TRY(batch_inverse_mont<P>(ctx, m, nullptr, n, 1, (Fp<P> *)out, &zero, st));
    if (zero) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_batch_inverse: an element is zero");
    return ZKB_OK;
}

template <class P>
static int perm_grand_product_t(zkb_ctx *ctx, uint64_t n, uint32_t ncols, const void *cols, const void *sid, const void *ssigma,
                                const uint32_t *beta, const uint32_t *gamma, void *v_out, cudaStream_t st) {
    typedef Fp<P> F;
    F b, g;
    memcpy(b.l, beta, sizeof(b.l));
    memcpy(g.l, gamma, sizeof(g.l));
    for (const F *x : {&b, &g}) {
        bool lt = false;
        for (int i = F::N - 1; i >= 0 && !lt; i--) {
            if (x->l[i] < P::mod(i)) lt = true;
            else if (x->l[i] > P::mod(i)) break;
        }
        if (!lt) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "beta / gamma >= modulus");
    }
    void *p;
    ZKB_TRY(ctx_scratch(ctx, "perm_rows", (size_t)3 * n * sizeof(F), &p));
    F *nom = (F *)p, *den = nom + n, *ratio = den + n;
    // beta in Montgomery form: canonical sid * mont(beta) = canonical sid beta; gamma stays canonical
    perm_rows_kernel<P><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, ncols, (const F *)cols, (const F *)sid, (const F *)ssigma,
                                                                    b.to_mont(), g, nom, den);
    ctx->launches++;
    ZKB_CUDA_OK(ctx, cudaGetLastError());
    bool zero = false;
    ZKB_TRY(batch_inverse_mont<P>(ctx, den, nom, n, 0, ratio, &zero, st));
    if (zero) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "permutation grand product: a denominator is zero");
    // V_P[0] = 1, V_P[j] = prod_{k < j} ratio[k]
    return scan_run<P>(ctx, ratio, n, 0, 1, 1, (F *)v_out, nullptr, nullptr, st);
}

template <class P>
static bool scan_canonical(const Fp<P> &x) {
    for (int i = Fp<P>::N - 1; i >= 0; i--) {
        if (x.l[i] < P::mod(i)) return true;
        if (x.l[i] > P::mod(i)) return false;
    }
    return false;
}

template <class P>
static int lookup_grand_product_t(zkb_ctx *ctx, uint64_t n, uint64_t usable, uint32_t n_in, const void *inputs, uint32_t n_val,
                                  const void *values, uint32_t n_sorted, const void *sorted, const uint32_t *beta,
                                  const uint32_t *gamma, void *v_out, cudaStream_t st) {
    typedef Fp<P> F;
    F b, g;
    memcpy(b.l, beta, sizeof(b.l));
    memcpy(g.l, gamma, sizeof(g.l));
    if (!scan_canonical<P>(b) || !scan_canonical<P>(g)) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "beta / gamma >= modulus");
    const F bm = b.to_mont(), one_beta = F::one() + bm;               // Montgomery
    const F part1 = (one_beta * g.to_mont()).from_mont();             // canonical (1 + beta) gamma
    const F pw = one_beta.pow_u64(n_in);                              // Montgomery (1 + beta)^n_in
    void *p;
    ZKB_TRY(ctx_scratch(ctx, "perm_rows", (size_t)3 * n * sizeof(F), &p));
    F *nom = (F *)p, *den = nom + n, *ratio = den + n;
    lookup_rows_kernel<P><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(n, usable, n_in, (const F *)inputs, n_val, (const F *)values, n_sorted,
                                                                      (const F *)sorted, bm, g, part1, pw, nom, den);
    ctx->launches++;
    ZKB_CUDA_OK(ctx, cudaGetLastError());
    bool zero = false;
    ZKB_TRY(batch_inverse_mont<P>(ctx, den, nom, n, 0, ratio, &zero, st));
    if (zero) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "lookup grand product: a denominator is zero");
    ZKB_TRY(scan_run<P>(ctx, ratio, n, 0, 1, 1, (F *)v_out, nullptr, nullptr, st));
    // V_L[k] stays zero beyond the usable rows (the polynomial is created zero-filled, :382-383)
    if (usable + 1 < n) ZKB_CUDA_OK(ctx, cudaMemsetAsync((F *)v_out + usable + 1, 0, (size_t)(n - usable - 1) * sizeof(F), st));
    return ZKB_OK;
}

#define ZKB_DISPATCH_SCAN_FIELD(field, FN, ...)                                  \
    switch (field) {                                                             \
        case ZKB_FIELD_BLS12_381_FR: return FN<params::Bls12381Fr>(__VA_ARGS__); \
        case ZKB_FIELD_BN254_FR: return FN<params::Bn254Fr>(__VA_ARGS__);        \
        case ZKB_FIELD_PALLAS_FP: return FN<params::PallasFp>(__VA_ARGS__);      \
        case ZKB_FIELD_PALLAS_FQ: return FN<params::PallasFq>(__VA_ARGS__);      \
        default: return ZKB_ERR_INVALID_ARGUMENT;                                \
    }

static int prefix_product_dispatch(zkb_ctx *ctx, int field, uint64_t n, const void *in, void *out, int exclusive, cudaStream_t st) {
    ZKB_DISPATCH_SCAN_FIELD(field, prefix_product_t, ctx, n, in, out, exclusive, st)
}
static int batch_inverse_dispatch(zkb_ctx *ctx, int field, uint64_t n, const void *in, void *out, cudaStream_t st) {
    ZKB_DISPATCH_SCAN_FIELD(field, batch_inverse_t, ctx, n, in, out, st)
}
static int perm_dispatch(zkb_ctx *ctx, int field, uint64_t n, uint32_t ncols, const void *cols, const void *sid, const void *ssigma,
                         const uint32_t *beta, const uint32_t *gamma, void *v_out, cudaStream_t st) {
    ZKB_DISPATCH_SCAN_FIELD(field, perm_grand_product_t, ctx, n, ncols, cols, sid, ssigma, beta, gamma, v_out, st)
}

static int lookup_dispatch(zkb_ctx *ctx, int field, uint64_t n, uint64_t usable, uint32_t n_in, const void *inputs, uint32_t n_val,
                           const void *values, uint32_t n_sorted, const void *sorted, const uint32_t *beta, const uint32_t *gamma,
                           void *v_out, cudaStream_t st) {
    ZKB_DISPATCH_SCAN_FIELD(field, lookup_grand_product_t, ctx, n, usable, n_in, inputs, n_val, values, n_sorted, sorted, beta, gamma, v_out, st)
}

extern "C" {

int zkb_lookup_grand_product(zkb_ctx *ctx, int field, uint64_t n, uint64_t usable_rows, uint32_t n_inputs, const void *inputs_device,
                             uint32_t n_values, const void *values_device, uint32_t n_sorted, const void *sorted_device,
                             const uint32_t *beta, const uint32_t *gamma, void *v_out_device, void *stream) {
    if (!ctx) return ZKB_ERR_INVALID_ARGUMENT;
    if (field < ZKB_FIELD_BLS12_381_FR || field > ZKB_FIELD_PALLAS_FQ || n == 0 || usable_rows >= n || !beta || !gamma || !v_out_device ||
        (n_inputs && !inputs_device) || (n_values && !values_device) || (n_sorted && !sorted_device))
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_lookup_grand_product: bad arguments (usable_rows must be < n)");
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    return lookup_dispatch(ctx, field, n, usable_rows, n_inputs, inputs_device, n_values, values_device, n_sorted, sorted_device, beta, gamma,
                           v_out_device, (cudaStream_t)stream);
}

int zkb_prefix_product(zkb_ctx *ctx, int field, uint64_t n, const void *in_device, void *out_device, int exclusive, void *stream) {
    if (!ctx) return ZKB_ERR_INVALID_ARGUMENT;
    if (field < ZKB_FIELD_BLS12_381_FR || field > ZKB_FIELD_PALLAS_FQ || (n && (!in_device || !out_device)))
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_prefix_product: bad arguments");
    if (n == 0) return ZKB_OK;
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    return prefix_product_dispatch(ctx, field, n, in_device, out_device, exclusive, (cudaStream_t)stream);
}

int zkb_batch_inverse(zkb_ctx *ctx, int field, uint64_t n, const void *in_device, void *out_device, void *stream) {
    if (!ctx) return ZKB_ERR_INVALID_ARGUMENT;
    if (field < ZKB_FIELD_BLS12_381_FR || field > ZKB_FIELD_PALLAS_FQ || (n && (!in_device || !out_device)))
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_batch_inverse: bad arguments");
    if (n == 0) return ZKB_OK;
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    return batch_inverse_dispatch(ctx, field, n, in_device, out_device, (cudaStream_t)stream);
}

int zkb_permutation_grand_product(zkb_ctx *ctx, int field, uint64_t n, uint32_t ncols, const void *columns_device,
                                  const void *s_id_device, const void *s_sigma_device, const uint32_t *beta, const uint32_t *gamma,
                                  void *v_out_device, void *stream) {
    if (!ctx) return ZKB_ERR_INVALID_ARGUMENT;
    if (field < ZKB_FIELD_BLS12_381_FR || field > ZKB_FIELD_PALLAS_FQ || n == 0 || ncols == 0 || !columns_device || !s_id_device ||
        !s_sigma_device || !beta || !gamma || !v_out_device)
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_permutation_grand_product: bad arguments");
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    return perm_dispatch(ctx, field, n, ncols, columns_device, s_id_device, s_sigma_device, beta, gamma, v_out_device, (cudaStream_t)stream);
}

}  // extern "C"
