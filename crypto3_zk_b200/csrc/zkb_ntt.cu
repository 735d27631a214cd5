// NTT / LDE / pointwise / FRI-fold runtime of libzkb200.so: table caches, pass launches, C ABI.
// Kernel bodies are in zkb_ntt_pass.cuh; the plan and table definitions in zkb_ntt_plan.h /
// zkb_ntt_tables.cuh.  See include/zkb200.h for the reference call sites each entry point serves.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <atomic>
#include "zkb_internal.h"
#include "zkb_ntt_plan.h"
#include "zkb_ntt_tables.cuh"

using namespace zkb;

// ------------------------------------------------------------------------------------ table kernels
template <class P>
__global__ void __launch_bounds__(256) powtab_kernel(Fp<P> base, Fp<P> scale, int two_d, int log_m, uint64_t count,
                                                     Fp<P> *out) {
    uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx < count) out[idx] = powtab_value<Fp<P>>(base, scale, two_d, log_m, idx);
}

template <class P>
static int fill_powtab(zkb_ctx *ctx, const Fp<P> &base, const Fp<P> &scale, int two_d, int log_m, uint64_t count,
                       void *out, cudaStream_t st) {
    uint64_t blocks = (count + 255) / 256;
    powtab_kernel<P><<<(unsigned)blocks, 256, 0, st>>>(base, scale, two_d, log_m, count, (Fp<P> *)out);
    ctx->launches++;
    ZKB_CUDA_OK(ctx, cudaGetLastError());
    return ZKB_OK;
}

static std::string hexkey(const uint32_t *l, int n) {
    char buf[16];
    std::string s;
    for (int i = n - 1; i >= 0; i--) {
        snprintf(buf, sizeof buf, "%08x", l[i]);
        s += buf;
    }
    return s;
}

// Collects (building on first use) every table one transform needs.
template <class P>
static int ntt_tables(zkb_ctx *ctx, const NttPlan &pl, int inverse, const uint32_t *shift, NttTables *tb,
                      cudaStream_t st) {
    typedef Fp<P> F;
    const int log_n = pl.log_n;
    const uint64_t N = 1ull << log_n;
    memset(tb, 0, sizeof(*tb));
    char key[160];
    bool created;
    void *p;
    // master in-tile twiddles
    snprintf(key, sizeof key, "tw:%d:%d", P::ID, inverse);
    ZKB_TRY(ctx_table(ctx, key, sizeof(F) << (ZKB_NTT_TW_LOG - 1), &p, &created));
    if (created) ZKB_TRY(fill_powtab<P>(ctx, ntt_omega<F, P>(ZKB_NTT_TW_LOG, inverse), F::one(), 0, 0,
                                        1ull << (ZKB_NTT_TW_LOG - 1), p, st));
    tb->tw = p;
    {
        const char *v = getenv("ZKB_NTT_TW_SMEM");      // A/B switch for profiles/; default on
        tb->tw_in_smem = v && *v ? atoi(v) : 1;
    }
    // inter-pass twiddles
    F ninv = ntt_n_inv<F>(log_n);
    for (int i = 0; i + 1 < pl.n_passes; i++) {
        int lmp = pl.log_m(i), lm = pl.log_m(i + 1);
        bool scaled = (i == 0 && inverse);
        snprintf(key, sizeof key, "inter:%d:%d:%d:%d:%d", P::ID, inverse, lmp, lm, scaled ? log_n : -1);
        ZKB_TRY(ctx_table(ctx, key, sizeof(F) << lmp, &p, &created));
        if (created)
            ZKB_TRY(fill_powtab<P>(ctx, ntt_omega<F, P>(lmp, inverse), scaled ? ninv : F::one(), 1, lm, 1ull << lmp, p, st));
        tb->inter[i] = p;
    }
    if (shift) {
        F g;
        memcpy(g.l, shift, sizeof(g.l));
        for (int i = F::N - 1; i >= 0; i--) {  // must be canonical
            if (g.l[i] < P::mod(i)) break;
            if (g.l[i] > P::mod(i) || i == 0) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "coset shift >= modulus");
        }
        if (g.is_zero()) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "coset shift is zero");
        bool scaled = inverse && pl.n_passes == 1;
        snprintf(key, sizeof key, "coset:%d:%d:%d:%d:%s", P::ID, log_n, inverse, (int)scaled, hexkey(shift, F::N).c_str());
        ZKB_TRY(ctx_table(ctx, key, sizeof(F) << log_n, &p, &created));
        if (created) {
            F gm = g.to_mont();
            if (inverse) gm = gm.inverse();
            ZKB_TRY(fill_powtab<P>(ctx, gm, scaled ? ninv : F::one(), 0, 0, N, p, st));
        }
        if (!inverse) { tb->load_tab = p; tb->load_mask = N - 1; }
        else { tb->store_tab = p; tb->store_mask = N - 1; }
    } else if (inverse && pl.n_passes == 1) {
        snprintf(key, sizeof key, "ninv:%d:%d", P::ID, log_n);
        ZKB_TRY(ctx_table(ctx, key, sizeof(F), &p, &created));
        if (created) ZKB_CUDA_OK(ctx, cudaMemcpyAsync(p, &ninv, sizeof(F), cudaMemcpyHostToDevice, st));
        tb->store_tab = p;
        tb->store_mask = 0;
    }
    return ZKB_OK;
}

template <class P>
static int ntt_launch(zkb_ctx *ctx, const NttPassParams &q, cudaStream_t st) {
    // function attributes are per DEVICE: one bit per device id (a process-wide flag left every device but the first
    // without the opt-in shared-memory size - zkb_multi with four or more devices, found by the 8-GPU run of round 2)
    static std::atomic<uint64_t> attr_set{0};
    const uint64_t bit = 1ull << (ctx->device & 63);
    if (!(attr_set.load(std::memory_order_acquire) & bit)) {
        ZKB_CUDA_OK(ctx, cudaFuncSetAttribute(ntt_pass_kernel<P>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              (int)ntt_smem_bytes(ZKB_NTT_MAX_LOG_R)));
        attr_set.fetch_or(bit, std::memory_order_release);
    }
    uint64_t tiles = ntt_pass_tiles(q);
    if (tiles == 0) return ZKB_OK;
    if (tiles > 0x7fffffffull) return ctx_fail(ctx, ZKB_ERR_UNSUPPORTED, "too many tiles for one launch");
    ntt_pass_kernel<P><<<(unsigned)tiles, ZKB_NTT_THREADS, ntt_smem_bytes(q.log_r), st>>>(q);
    ctx->launches++;
    ZKB_CUDA_OK(ctx, cudaGetLastError());
    return ZKB_OK;
}

// the known outputs of an LDE (see NttPassParams::known_log): a strided copy of the input evaluations
__global__ void __launch_bounds__(256) ntt_known_scatter_kernel(const u128 *__restrict__ src, u128 *__restrict__ dst, uint64_t total,
                                                                int log_n, int known_log, uint64_t known_poly_stride,
                                                                uint64_t out_poly_stride) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * total) return;
    uint64_t s, d;
    ntt_known_scatter_index(i >> 1, log_n, known_log, known_poly_stride, out_poly_stride, &s, &d);
    dst[2 * d + (i & 1)] = src[2 * s + (i & 1)];
}

template <class P>
static int ntt_device_t(zkb_ctx *ctx, int log_n, uint32_t batch, const void *d_in, void *d_out, int inverse,
                        const uint32_t *shift, uint64_t in_poly_stride, uint64_t in_valid, cudaStream_t st,
                        const void *known_src, int known_log, uint64_t known_poly_stride) {
    typedef Fp<P> F;
    if (log_n > P::TWO_ADICITY) return ctx_fail(ctx, ZKB_ERR_DOMAIN_TOO_LARGE, "2^log_n exceeds the two-adicity");
    const uint64_t N = 1ull << log_n;
    NttPlan pl = ntt_make_plan(log_n, in_valid < N);
    if (pl.n_passes < 1) pl = ntt_make_plan(log_n);
    if (pl.n_passes < 1) return ctx_fail(ctx, ZKB_ERR_UNSUPPORTED, "no pass plan");
    NttTables tb;
    ZKB_TRY(ntt_tables<P>(ctx, pl, inverse, shift, &tb, st));
    // chunk the batch so the work buffer respects the scratch limit
    uint64_t poly_bytes = N * sizeof(F);
    uint32_t chunk = batch;
    if (pl.n_passes > 1) {
        uint64_t fit = ctx->scratch_limit / poly_bytes;
        if (fit < 1) fit = 1;
        if (fit < chunk) chunk = (uint32_t)fit;
    }
    void *work = nullptr;
    if (pl.n_passes > 1) ZKB_TRY(ctx_scratch(ctx, "ntt_work", (size_t)chunk * poly_bytes, &work));
    for (uint32_t b0 = 0; b0 < batch; b0 += chunk) {
        uint32_t nb = batch - b0 < chunk ? batch - b0 : chunk;
        const char *cin = (const char *)d_in + (size_t)b0 * in_poly_stride * sizeof(F);
        char *cout = (char *)d_out + (size_t)b0 * poly_bytes;
        const char *ksrc = known_src ? (const char *)known_src + (size_t)b0 * known_poly_stride * sizeof(F) : nullptr;
        auto passes = ntt_build_passes(pl, tb, cin, cout, work, nb, in_poly_stride, N, in_valid, ksrc, known_log, known_poly_stride);
        for (auto &q : passes) ZKB_TRY(ntt_launch<P>(ctx, q, st));
        if (passes.back().known_log > 0) {
            const int kl = passes.back().known_log;
            const uint64_t total = (uint64_t)nb << (log_n - kl);
            ntt_known_scatter_kernel<<<(unsigned)((2 * total + 255) / 256), 256, 0, st>>>((const u128 *)ksrc, (u128 *)cout, total, log_n, kl,
                                                                                       known_poly_stride, N);
            ctx->launches++;
            ZKB_CUDA_OK(ctx, cudaGetLastError());
        }
    }
    return ZKB_OK;
}

#define ZKB_DISPATCH_NTT_FIELD(field, FN, ...)                                   \
    switch (field) {                                                             \
        case ZKB_FIELD_BLS12_381_FR: return FN<params::Bls12381Fr>(__VA_ARGS__); \
        case ZKB_FIELD_BN254_FR: return FN<params::Bn254Fr>(__VA_ARGS__);        \
        case ZKB_FIELD_PALLAS_FP: return FN<params::PallasFp>(__VA_ARGS__);      \
        case ZKB_FIELD_PALLAS_FQ: return FN<params::PallasFq>(__VA_ARGS__);      \
        default: return ZKB_ERR_INVALID_ARGUMENT;                                \
    }

namespace zkb {

int ntt_device(zkb_ctx *ctx, int field, int log_n, uint32_t batch, const void *d_in, void *d_out, int inverse,
               const uint32_t *coset_shift, uint64_t in_poly_stride, uint64_t in_valid_elems, cudaStream_t st,
               const void *known_src, int known_log, uint64_t known_poly_stride) {
    if (batch == 0) return ZKB_OK;
    if (log_n == 0) {  // size-1 transform is the identity (a coset shift g^0 = 1, 1/n = 1)
        if (d_in != d_out)
            ZKB_CUDA_OK(ctx, cudaMemcpy2DAsync(d_out, 32, d_in, in_poly_stride * 32, 32, batch, cudaMemcpyDeviceToDevice, st));
        return ZKB_OK;
    }
    ZKB_DISPATCH_NTT_FIELD(field, ntt_device_t, ctx, log_n, batch, d_in, d_out, inverse, coset_shift, in_poly_stride,
                           in_valid_elems, st, known_src, known_log, known_poly_stride)
}

int lde_device(zkb_ctx *ctx, int field, int log_n_in, int log_n_out, uint32_t batch, const void *d_in, void *d_out,
               cudaStream_t st, void *d_coef_out) {
    if (batch == 0) return ZKB_OK;
    const uint64_t Nin = 1ull << log_n_in, Nout = 1ull << log_n_out;
    if (log_n_in == log_n_out) {
        if (d_in != d_out) ZKB_CUDA_OK(ctx, cudaMemcpyAsync(d_out, d_in, (size_t)batch * Nin * 32, cudaMemcpyDeviceToDevice, st));
        if (d_coef_out) ZKB_TRY(ntt_device(ctx, field, log_n_in, batch, d_in, d_coef_out, 1, nullptr, Nin, Nin, st));
        return ZKB_OK;
    }
    // coefficients of a chunk of polynomials, then the zero-padded forward transform
    uint64_t per_poly = (Nin + Nout) * 32;   // coefficient buffer + ntt work buffer
    uint32_t chunk = batch;
    uint64_t fit = ctx->scratch_limit / per_poly;
    if (fit < 1) fit = 1;
    if (fit < chunk) chunk = (uint32_t)fit;
    void *coef_scratch = nullptr;
    if (!d_coef_out) ZKB_TRY(ctx_scratch(ctx, "lde_coef", (size_t)chunk * Nin * 32, &coef_scratch));
    for (uint32_t b0 = 0; b0 < batch; b0 += chunk) {
        uint32_t nb = batch - b0 < chunk ? batch - b0 : chunk;
        const char *cin = (const char *)d_in + (size_t)b0 * Nin * 32;
        char *cout = (char *)d_out + (size_t)b0 * Nout * 32;
        void *coef = d_coef_out ? (void *)((char *)d_coef_out + (size_t)b0 * Nin * 32) : coef_scratch;
        ZKB_TRY(ntt_device(ctx, field, log_n_in, nb, cin, coef, 1, nullptr, Nin, Nin, st));
        // out[8 i] (blow-up 8) are the input evaluations: the forward transform neither computes nor stores them
        ZKB_TRY(ntt_device(ctx, field, log_n_out, nb, coef, cout, 0, nullptr, Nin, Nin, st, cin, log_n_out - log_n_in, Nin));
    }
    return ZKB_OK;
}

}  // namespace zkb

// ------------------------------------------------------------------------------------ pointwise ops
template <class P>
__global__ void __launch_bounds__(256) vec_kernel(int op, uint64_t n, const Fp<P> *a, const Fp<P> *b, const Fp<P> *c,
                                                  Fp<P> s_mont, Fp<P> *out) {
    typedef Fp<P> F;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    F x = a[i], y = b[i], r;
    switch (op) {
        case ZKB_VEC_MUL: r = (x * y) * F::r2(); break;            // canonical in/out: (xy/R) * R^2 / R
        case ZKB_VEC_SUB: r = x - y; break;
        case ZKB_VEC_ADD: r = x + y; break;
        default: {                                                  // (a*b - c) * s
            F ab = (x * y) * F::r2();
            r = (ab - c[i]) * s_mont;
        }
    }
    out[i] = r;
}

template <class P>
static int vec_t(zkb_ctx *ctx, int op, uint64_t n, const void *a, const void *b, const void *c, const uint32_t *scalar,
                 void *out, cudaStream_t st) {
    typedef Fp<P> F;
    F s = F::one();
    if (op == ZKB_VEC_MUL_SUB_SCALE) {
        if (!scalar || !c) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "MUL_SUB_SCALE needs c and scalar");
        memcpy(s.l, scalar, sizeof(s.l));
        s = s.to_mont();
    }
    uint64_t blocks = (n + 255) / 256;
    if (blocks == 0) return ZKB_OK;
    vec_kernel<P><<<(unsigned)blocks, 256, 0, st>>>(op, n, (const F *)a, (const F *)b, (const F *)c, s, (F *)out);
    ctx->launches++;
    ZKB_CUDA_OK(ctx, cudaGetLastError());
    return ZKB_OK;
}

// ------------------------------------------------------------------------------------ FRI fold
// out[i] = 1/2 ((1 + alpha w^-i) f[i] + (1 - alpha w^-i) f[i + n/2]); w^-i comes from a cached power
// table instead of the reference's serial acc *= w^-1 (fold_polynomial.hpp:87-90).
template <class P>
__global__ void __launch_bounds__(256) fold_kernel(uint64_t half, const Fp<P> *f, const Fp<P> *winv_pow, Fp<P> alpha_mont,
                                                   Fp<P> *out) {
    typedef Fp<P> F;
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= half) return;
    F acc = alpha_mont * winv_pow[i];                 // alpha * w^-i, Montgomery form
    F a = f[i], b = f[i + half];                      // canonical
    // 1/2 ((a + b) + acc (a - b)) ; mont_mul(canonical, montgomery) = canonical product
    F s = a + b, d = (a - b) * acc;
    out[i] = (s + d) * F::two_inv();
}

template <class P>
static int fold_t(zkb_ctx *ctx, int log_n, const void *f, const uint32_t *alpha, void *out, cudaStream_t st) {
    typedef Fp<P> F;
    if (log_n < 1) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "fold needs log_n >= 1");
    if (log_n > P::TWO_ADICITY) return ctx_fail(ctx, ZKB_ERR_DOMAIN_TOO_LARGE, "2^log_n exceeds the two-adicity");
    uint64_t half = 1ull << (log_n - 1);
    char key[64];
    snprintf(key, sizeof key, "foldw:%d:%d", P::ID, log_n);
    void *tab;
    bool created;
    ZKB_TRY(ctx_table(ctx, key, sizeof(F) * half, &tab, &created));
    if (created) ZKB_TRY(fill_powtab<P>(ctx, ntt_omega<F, P>(log_n, true), F::one(), 0, 0, half, tab, st));
    F a;
    memcpy(a.l, alpha, sizeof(a.l));
    a = a.to_mont();
    uint64_t blocks = (half + 255) / 256;
    fold_kernel<P><<<(unsigned)blocks, 256, 0, st>>>(half, (const F *)f, (const F *)tab, a, (F *)out);
    ctx->launches++;
    ZKB_CUDA_OK(ctx, cudaGetLastError());
    return ZKB_OK;
}

namespace zkb {
int fold_device(zkb_ctx *ctx, int field, int log_n, const void *d_f, const uint32_t *alpha, void *d_out, cudaStream_t st) {
    ZKB_DISPATCH_NTT_FIELD(field, fold_t, ctx, log_n, d_f, alpha, d_out, st)
}
}  // namespace zkb

// ------------------------------------------------------------------------------------ C ABI
namespace {
struct Staged {  // host<->device staging for ZKB_MEM_HOST callers
    zkb_ctx *ctx;
    cudaStream_t st;
    int mem;
    Staged(zkb_ctx *c, cudaStream_t s, int m) : ctx(c), st(s), mem(m) {}
    int in(const char *role, const void *src, size_t bytes, const void **dev) {
        if (mem == ZKB_MEM_DEVICE || !src) { *dev = src; return ZKB_OK; }
        void *d;
        ZKB_TRY(ctx_scratch(ctx, role, bytes, &d));
        ZKB_CUDA_OK(ctx, cudaMemcpyAsync(d, src, bytes, cudaMemcpyHostToDevice, st));
        *dev = d;
        return ZKB_OK;
    }
    int out_buf(const char *role, void *dst, size_t bytes, void **dev) {
        if (mem == ZKB_MEM_DEVICE) { *dev = dst; return ZKB_OK; }
        return ctx_scratch(ctx, role, bytes, dev);
    }
    int out(void *dst, const void *dev, size_t bytes) {
        if (mem == ZKB_MEM_DEVICE) return ZKB_OK;
        ZKB_CUDA_OK(ctx, cudaMemcpyAsync(dst, dev, bytes, cudaMemcpyDeviceToHost, st));
        ZKB_CUDA_OK(ctx, cudaStreamSynchronize(st));
        return ZKB_OK;
    }
};
bool is_ntt_field(int f) { return f >= ZKB_FIELD_BLS12_381_FR && f <= ZKB_FIELD_PALLAS_FQ; }

// Host-buffer callers (ZKB_MEM_HOST) of the batched transforms: polynomials stream through two device
// slots so that the upload of chunk k+1 and the download of chunk k-1 overlap the kernels of chunk k
// (PCIe is full duplex; the LDE of config #2 returns 8x what it reads, so the download is the bound).
// `compute(nb, din, dout)` enqueues the kernels for nb polynomials on `st`.
template <class Fn>
int host_pipeline(zkb_ctx *ctx, cudaStream_t st, uint32_t batch, size_t in_poly_bytes, size_t out_poly_bytes,
                  const void *in, void *out, Fn compute) {
    if (!ctx->copy_in) {
        ZKB_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
        ZKB_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            ZKB_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->ev_h2d[i], cudaEventDisableTiming));
            ZKB_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->ev_comp[i], cudaEventDisableTiming));
            ZKB_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->ev_d2h[i], cudaEventDisableTiming));
        }
    }
    const size_t target = 256ull << 20;   // bytes of output per chunk
    uint32_t chunk = (uint32_t)(target / out_poly_bytes);
    if (chunk < 1) chunk = 1;
    if (chunk > batch) chunk = batch;
    void *din[2], *dout[2];
    char *base_in, *base_out;
    ZKB_TRY(ctx_scratch(ctx, "pipe_in", 2 * chunk * in_poly_bytes, (void **)&base_in));
    ZKB_TRY(ctx_scratch(ctx, "pipe_out", 2 * chunk * out_poly_bytes, (void **)&base_out));
    for (int i = 0; i < 2; i++) {
        din[i] = base_in + (size_t)i * chunk * in_poly_bytes;
        dout[i] = base_out + (size_t)i * chunk * out_poly_bytes;
    }
    int status = ZKB_OK;
    uint32_t k = 0;
    for (uint32_t b0 = 0; b0 < batch && status == ZKB_OK; b0 += chunk, k++) {
        const uint32_t nb = batch - b0 < chunk ? batch - b0 : chunk;
        const int s = k & 1;
        cudaError_t e = cudaSuccess;
        if (k >= 2) e = cudaStreamWaitEvent(ctx->copy_in, ctx->ev_comp[s], 0);     // slot's previous kernels have read it
        if (e == cudaSuccess)
            e = cudaMemcpyAsync(din[s], (const char *)in + (size_t)b0 * in_poly_bytes, nb * in_poly_bytes,
                                cudaMemcpyHostToDevice, ctx->copy_in);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_h2d[s], ctx->copy_in);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st, ctx->ev_h2d[s], 0);
        if (e == cudaSuccess && k >= 2) e = cudaStreamWaitEvent(st, ctx->ev_d2h[s], 0);   // slot's previous output is home
        if (e != cudaSuccess) { status = ctx_fail(ctx, ZKB_ERR_CUDA, std::string("host pipeline: ") + cudaGetErrorString(e)); break; }
        status = compute(nb, (const void *)din[s], dout[s]);
        if (status != ZKB_OK) break;
        e = cudaEventRecord(ctx->ev_comp[s], st);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copy_out, ctx->ev_comp[s], 0);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync((char *)out + (size_t)b0 * out_poly_bytes, dout[s], nb * out_poly_bytes,
                                cudaMemcpyDeviceToHost, ctx->copy_out);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_d2h[s], ctx->copy_out);
        if (e != cudaSuccess) status = ctx_fail(ctx, ZKB_ERR_CUDA, std::string("host pipeline: ") + cudaGetErrorString(e));
    }
    // the call returns with the results in `out` (and nothing in flight, also on the error path)
    cudaError_t e1 = cudaStreamSynchronize(ctx->copy_in), e2 = cudaStreamSynchronize(st), e3 = cudaStreamSynchronize(ctx->copy_out);
    if (status != ZKB_OK) return status;
    ZKB_CUDA_OK(ctx, e1);
    ZKB_CUDA_OK(ctx, e2);
    ZKB_CUDA_OK(ctx, e3);
    return ZKB_OK;
}
}  // namespace

extern "C" {

int zkb_ntt(zkb_ctx *ctx, int field, int log_n, uint32_t batch, const void *in, void *out, int inverse,
            const uint32_t *coset_shift, int mem, void *stream) {
    if (!ctx) return ZKB_ERR_INVALID_ARGUMENT;
    if (!is_ntt_field(field) || log_n < 0 || (batch && (!in || !out)))
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_ntt: bad field/log_n/pointers");
    if (log_n > zkb_field_two_adicity(field)) return ctx_fail(ctx, ZKB_ERR_DOMAIN_TOO_LARGE, "2^log_n exceeds the two-adicity");
    if (batch == 0) return ZKB_OK;
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    size_t bytes = ((size_t)batch << log_n) * 32;
    if (mem != ZKB_MEM_DEVICE)
        return host_pipeline(ctx, st, batch, bytes / batch, bytes / batch, in, out, [&](uint32_t nb, const void *di, void *dout) {
            return ntt_device(ctx, field, log_n, nb, di, dout, inverse, coset_shift, 1ull << log_n, 1ull << log_n, st);
        });
    return ntt_device(ctx, field, log_n, batch, in, out, inverse, coset_shift, 1ull << log_n, 1ull << log_n, st);
}

int zkb_lde(zkb_ctx *ctx, int field, int log_n_in, int log_n_out, uint32_t batch, const void *in, void *out, int mem,
            void *stream) {
    if (!ctx) return ZKB_ERR_INVALID_ARGUMENT;
    if (!is_ntt_field(field) || log_n_in < 1 || log_n_out < log_n_in || (batch && (!in || !out)))
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_lde: need 1 <= log_n_in <= log_n_out");
    if (log_n_out > zkb_field_two_adicity(field)) return ctx_fail(ctx, ZKB_ERR_DOMAIN_TOO_LARGE, "2^log_n_out exceeds the two-adicity");
    if (batch == 0) return ZKB_OK;
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    if (mem != ZKB_MEM_DEVICE)
        return host_pipeline(ctx, st, batch, (size_t)32 << log_n_in, (size_t)32 << log_n_out, in, out,
                             [&](uint32_t nb, const void *di, void *dout) {
                                 return lde_device(ctx, field, log_n_in, log_n_out, nb, di, dout, st);
                             });
    return lde_device(ctx, field, log_n_in, log_n_out, batch, in, out, st);
}

int zkb_lde_with_coefficients(zkb_ctx *ctx, int field, int log_n_in, int log_n_out, uint32_t batch, const void *in_device,
                              void *out_device, void *coefficients_out_device, void *stream) {
    if (!ctx) return ZKB_ERR_INVALID_ARGUMENT;
    if (!is_ntt_field(field) || log_n_in < 1 || log_n_out < log_n_in || (batch && (!in_device || !out_device || !coefficients_out_device)))
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_lde_with_coefficients: need 1 <= log_n_in <= log_n_out and device buffers");
    if (log_n_out > zkb_field_two_adicity(field)) return ctx_fail(ctx, ZKB_ERR_DOMAIN_TOO_LARGE, "2^log_n_out exceeds the two-adicity");
    if (batch == 0) return ZKB_OK;
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    return lde_device(ctx, field, log_n_in, log_n_out, batch, in_device, out_device, (cudaStream_t)stream, coefficients_out_device);
}

int zkb_vec(zkb_ctx *ctx, int field, int op, uint64_t n, const void *a, const void *b, const void *c,
            const uint32_t *scalar, void *out, int mem, void *stream) {
    if (!ctx) return ZKB_ERR_INVALID_ARGUMENT;
    if (!is_ntt_field(field) || op < 0 || op > ZKB_VEC_MUL_SUB_SCALE || (n && (!a || !b || !out)))
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_vec: bad arguments");
    if (n == 0) return ZKB_OK;
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    size_t bytes = (size_t)n * 32;
    Staged sg(ctx, st, mem);
    const void *da, *db, *dc;
    void *dout;
    ZKB_TRY(sg.in("vec_a", a, bytes, &da));
    ZKB_TRY(sg.in("vec_b", b, bytes, &db));
    ZKB_TRY(sg.in("vec_c", c, bytes, &dc));
    ZKB_TRY(sg.out_buf("vec_a", out, bytes, &dout));
    int s;
    switch (field) {
        case ZKB_FIELD_BLS12_381_FR: s = vec_t<params::Bls12381Fr>(ctx, op, n, da, db, dc, scalar, dout, st); break;
        case ZKB_FIELD_BN254_FR: s = vec_t<params::Bn254Fr>(ctx, op, n, da, db, dc, scalar, dout, st); break;
        case ZKB_FIELD_PALLAS_FP: s = vec_t<params::PallasFp>(ctx, op, n, da, db, dc, scalar, dout, st); break;
        default: s = vec_t<params::PallasFq>(ctx, op, n, da, db, dc, scalar, dout, st); break;
    }
    ZKB_TRY(s);
    return sg.out(out, dout, bytes);
}

int zkb_fri_fold(zkb_ctx *ctx, int field, int log_n, const void *f, const uint32_t *alpha, void *out, int mem,
                 void *stream) {
    if (!ctx) return ZKB_ERR_INVALID_ARGUMENT;
    if (!is_ntt_field(field) || !f || !alpha || !out || log_n < 1)
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_fri_fold: bad arguments");
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    size_t ib = ((size_t)1 << log_n) * 32, ob = ib / 2;
    Staged sg(ctx, st, mem);
    const void *df;
    void *dout;
    ZKB_TRY(sg.in("io_in", f, ib, &df));
    ZKB_TRY(sg.out_buf("io_out", out, ob, &dout));
    ZKB_TRY(fold_device(ctx, field, log_n, df, alpha, dout, st));
    return sg.out(out, dout, ob);
}

}  // extern "C"
