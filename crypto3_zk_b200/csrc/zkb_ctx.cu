// Context, field descriptors and the integer-pipe micro-benchmark of libzkb200.so.
#include <string.h>
#include "zkb_field.cuh"
#include "zkb_internal.h"
#include "zkb_ntt_tables.cuh"

using namespace zkb;

#include <mutex>
#include <set>

namespace zkb {

// Live contexts: handles (Merkle trees, sparse matrices) may outlive the context they were made with (Python frees them
// from __del__, after Context.close()); their free functions ask here before they touch the context's pools.
static std::mutex g_ctx_mutex;
static std::set<const zkb_ctx *> g_live_ctx;
bool ctx_alive(const zkb_ctx *ctx) {
    std::lock_guard<std::mutex> lk(g_ctx_mutex);
    return g_live_ctx.count(ctx) != 0;
}
static void ctx_register(const zkb_ctx *ctx, bool live) {
    std::lock_guard<std::mutex> lk(g_ctx_mutex);
    if (live) g_live_ctx.insert(ctx);
    else g_live_ctx.erase(ctx);
}

int ctx_fail(zkb_ctx *ctx, int status, const std::string &msg) {
    if (ctx) ctx->last_error = msg;
    return status;
}

int ctx_scratch(zkb_ctx *ctx, const char *role, size_t bytes, void **out) {
    zkb_ctx::Buf &b = ctx->scratch[role];
    if (b.cap < bytes) {
        if (b.p) cudaFree(b.p);
        b.p = nullptr;
        b.cap = 0;
        cudaError_t e = cudaMalloc(&b.p, bytes);
        if (e != cudaSuccess) {
            cudaGetLastError();
            return ctx_fail(ctx, ZKB_ERR_OUT_OF_MEMORY, std::string("cudaMalloc scratch ") + role + ": " + cudaGetErrorString(e));
        }
        b.cap = bytes;
    }
    *out = b.p;
    return ZKB_OK;
}

int ctx_tree_alloc(zkb_ctx *ctx, size_t bytes, void **out, size_t *cap) {
    // best fit among the pooled buffers that are not more than twice the request
    int best = -1;
    for (size_t i = 0; i < ctx->tree_pool.size(); i++) {
        size_t c = ctx->tree_pool[i].cap;
        if (c >= bytes && c <= 2 * bytes + 4096 && (best < 0 || c < ctx->tree_pool[best].cap)) best = (int)i;
    }
    if (best >= 0) {
        *out = ctx->tree_pool[best].p;
        *cap = ctx->tree_pool[best].cap;
        ctx->tree_pool_bytes -= *cap;
        ctx->tree_pool.erase(ctx->tree_pool.begin() + best);
        return ZKB_OK;
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {   // give the pooled buffers back to the driver and retry once
        cudaGetLastError();
        for (auto &b : ctx->tree_pool) cudaFree(b.p);
        ctx->tree_pool.clear();
        ctx->tree_pool_bytes = 0;
        e = cudaMalloc(&p, bytes);
    }
    if (e != cudaSuccess) {
        cudaGetLastError();
        return ctx_fail(ctx, ZKB_ERR_OUT_OF_MEMORY, std::string("cudaMalloc merkle tree: ") + cudaGetErrorString(e));
    }
    *out = p;
    *cap = bytes;
    return ZKB_OK;
}

void ctx_tree_release(zkb_ctx *ctx, void *p, size_t cap) {
    if (!p) return;
    if (ctx->tree_pool_bytes + cap > ctx->scratch_limit || ctx->tree_pool.size() >= 64) {
        cudaFree(p);
        return;
    }
    zkb_ctx::Buf b;
    b.p = p;
    b.cap = cap;
    ctx->tree_pool.push_back(b);
    ctx->tree_pool_bytes += cap;
}

int ctx_table(zkb_ctx *ctx, const std::string &key, size_t bytes, void **out, bool *created) {
    auto it = ctx->tables.find(key);
    if (it != ctx->tables.end()) {
        *out = it->second.p;
        *created = false;
        return ZKB_OK;
    }
    void *p = nullptr;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return ctx_fail(ctx, ZKB_ERR_OUT_OF_MEMORY, "cudaMalloc table " + key + ": " + cudaGetErrorString(e));
    }
    zkb_ctx::Buf b;
    b.p = p;
    b.cap = bytes;
    ctx->tables[key] = b;
    *out = p;
    *created = true;
    return ZKB_OK;
}

}  // namespace zkb

template <class P>
static int field_generator(uint32_t *out) {
    Fp<P> g = Fp<P>::generator().from_mont();
    memcpy(out, g.l, sizeof(g.l));
    return ZKB_OK;
}
template <class P>
static int field_unity_root(int log_n, uint32_t *out) {
    if (log_n < 0) return ZKB_ERR_INVALID_ARGUMENT;
    if (log_n > P::TWO_ADICITY) return ZKB_ERR_DOMAIN_TOO_LARGE;
    Fp<P> w = ntt_omega<Fp<P>, P>(log_n, false).from_mont();
    memcpy(out, w.l, sizeof(w.l));
    return ZKB_OK;
}

#define ZKB_DISPATCH_ANY_FIELD(field, FN, ...)                                   \
    switch (field) {                                                             \
        case ZKB_FIELD_BLS12_381_FR: return FN<params::Bls12381Fr>(__VA_ARGS__); \
        case ZKB_FIELD_BN254_FR: return FN<params::Bn254Fr>(__VA_ARGS__);        \
        case ZKB_FIELD_PALLAS_FP: return FN<params::PallasFp>(__VA_ARGS__);      \
        case ZKB_FIELD_PALLAS_FQ: return FN<params::PallasFq>(__VA_ARGS__);      \
        case ZKB_FIELD_BLS12_381_FQ: return FN<params::Bls12381Fq>(__VA_ARGS__); \
        case ZKB_FIELD_BN254_FQ: return FN<params::Bn254Fq>(__VA_ARGS__);        \
        default: return ZKB_ERR_INVALID_ARGUMENT;                                \
    }

extern "C" {

const char *zkb_version(void) { return "zkb200 0.1 (sm_100a)"; }

const char *zkb_status_string(int s) {
    switch (s) {
        case ZKB_OK: return "ok";
        case ZKB_ERR_INVALID_ARGUMENT: return "invalid argument";
        case ZKB_ERR_DOMAIN_TOO_LARGE: return "domain larger than the field's two-adicity";
        case ZKB_ERR_CUDA: return "CUDA error";
        case ZKB_ERR_OUT_OF_MEMORY: return "out of device memory";
        case ZKB_ERR_NO_DEVICE: return "no CUDA device (there is no CPU fallback)";
        case ZKB_ERR_UNSUPPORTED: return "unsupported";
    }
    return "unknown status";
}

int zkb_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

int zkb_ctx_create(int device, zkb_ctx **out) {
    if (!out) return ZKB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    int n = zkb_device_count();
    if (n <= 0) return ZKB_ERR_NO_DEVICE;
    if (device < 0 || device >= n) return ZKB_ERR_INVALID_ARGUMENT;
    if (cudaSetDevice(device) != cudaSuccess) return ZKB_ERR_CUDA;
    zkb_ctx *c = new zkb_ctx();
    c->device = device;
    cudaDeviceGetAttribute(&c->sm_count, cudaDevAttrMultiProcessorCount, device);
    zkb::ctx_register(c, true);
    *out = c;
    return ZKB_OK;
}

int zkb_ctx_release_caches(zkb_ctx *ctx) {
    if (!ctx) return ZKB_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto &kv : ctx->scratch)
        if (kv.second.p) cudaFree(kv.second.p);
    for (auto &kv : ctx->tables)
        if (kv.second.p) cudaFree(kv.second.p);
    for (auto &b : ctx->tree_pool)
        if (b.p) cudaFree(b.p);
    ctx->scratch.clear();
    ctx->tables.clear();
    ctx->tree_pool.clear();
    ctx->tree_pool_bytes = 0;
    return ZKB_OK;
}

void zkb_ctx_destroy(zkb_ctx *ctx) {
    if (!ctx) return;
    zkb::ctx_register(ctx, false);
    zkb_ctx_release_caches(ctx);
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    for (int i = 0; i < 2; i++) {
        if (ctx->ev_h2d[i]) cudaEventDestroy(ctx->ev_h2d[i]);
        if (ctx->ev_comp[i]) cudaEventDestroy(ctx->ev_comp[i]);
        if (ctx->ev_d2h[i]) cudaEventDestroy(ctx->ev_d2h[i]);
    }
    delete ctx;
}

const char *zkb_ctx_last_error(const zkb_ctx *ctx) { return ctx ? ctx->last_error.c_str() : ""; }

int zkb_ctx_set_scratch_limit(zkb_ctx *ctx, uint64_t bytes) {
    if (!ctx || bytes < (64ull << 20)) return ZKB_ERR_INVALID_ARGUMENT;
    ctx->scratch_limit = bytes;
    return ZKB_OK;
}

uint64_t zkb_ctx_kernel_launches(const zkb_ctx *ctx) { return ctx ? ctx->launches : 0; }

// ---------------------------------------------------------------------------- field descriptors
int zkb_field_limbs(int field) {
    switch (field) {
        case ZKB_FIELD_BLS12_381_FR: case ZKB_FIELD_BN254_FR: case ZKB_FIELD_PALLAS_FP:
        case ZKB_FIELD_PALLAS_FQ: case ZKB_FIELD_BN254_FQ: return 8;
        case ZKB_FIELD_BLS12_381_FQ: return 12;
    }
    return 0;
}

int zkb_field_two_adicity(int field) {
    switch (field) {
        case ZKB_FIELD_BLS12_381_FR: return params::Bls12381Fr::TWO_ADICITY;
        case ZKB_FIELD_BN254_FR: return params::Bn254Fr::TWO_ADICITY;
        case ZKB_FIELD_PALLAS_FP: return params::PallasFp::TWO_ADICITY;
        case ZKB_FIELD_PALLAS_FQ: return params::PallasFq::TWO_ADICITY;
        case ZKB_FIELD_BLS12_381_FQ: return params::Bls12381Fq::TWO_ADICITY;
        case ZKB_FIELD_BN254_FQ: return params::Bn254Fq::TWO_ADICITY;
    }
    return -1;
}

int zkb_field_generator(int field, uint32_t *out) {
    if (!out) return ZKB_ERR_INVALID_ARGUMENT;
    ZKB_DISPATCH_ANY_FIELD(field, field_generator, out)
}

int zkb_field_unity_root(int field, int log_n, uint32_t *out) {
    if (!out) return ZKB_ERR_INVALID_ARGUMENT;
    ZKB_DISPATCH_ANY_FIELD(field, field_unity_root, log_n, out)
}

}  // extern "C"

// ---------------------------------------------------------------------------- int-pipe roofline probe
// Four independent chains per thread of dependent Montgomery multiplications held in registers: this
// is the most favourable instruction mix the multiplier can see (no memory, full ILP), so
// field-mul/s measured here is the denominator for "fraction of integer-pipe peak".
template <class P>
__global__ void __launch_bounds__(256) bench_mul_kernel(Fp<P> seed, uint32_t iters, Fp<P> *out) {
    typedef Fp<P> F;
    F a = seed, b = seed, c = seed, d = seed;
    a.l[0] ^= threadIdx.x;
    b.l[0] ^= blockIdx.x;
    c.l[1] ^= threadIdx.x + 7;
    d.l[1] ^= blockIdx.x + 3;
    a = F::reduce_once(a); b = F::reduce_once(b); c = F::reduce_once(c); d = F::reduce_once(d);
    for (uint32_t i = 0; i < iters; i++) {
        a = a * b;
        b = b * c;
        c = c * d;
        d = d * a;
    }
    F r = a + b + c + d;
    if (r.l[0] == 0x12345678u && r.l[1] == 0x9abcdef0u) out[0] = r;  // practically never: keeps the chain live
}

template <class P>
static int bench_field_mul(zkb_ctx *ctx, uint32_t blocks, uint32_t threads, uint32_t iters, double *muls_per_s) {
    typedef Fp<P> F;
    void *out;
    ZKB_TRY(ctx_scratch(ctx, "bench", sizeof(F), &out));
    F seed = F::generator();
    cudaEvent_t e0, e1;
    ZKB_CUDA_OK(ctx, cudaEventCreate(&e0));
    ZKB_CUDA_OK(ctx, cudaEventCreate(&e1));
    bench_mul_kernel<P><<<blocks, threads>>>(seed, 16, (F *)out);  // warm-up
    ZKB_CUDA_OK(ctx, cudaEventRecord(e0));
    bench_mul_kernel<P><<<blocks, threads>>>(seed, iters, (F *)out);
    ZKB_CUDA_OK(ctx, cudaEventRecord(e1));
    ZKB_CUDA_OK(ctx, cudaEventSynchronize(e1));
    ctx->launches += 2;
    float ms = 0;
    ZKB_CUDA_OK(ctx, cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *muls_per_s = 4.0 * (double)iters * blocks * threads / (ms * 1e-3);
    return ZKB_OK;
}

extern "C" int zkb_bench_field_mul(zkb_ctx *ctx, int field, uint32_t blocks, uint32_t threads, uint32_t iters,
                                   double *muls_per_s) {
    if (!ctx || !muls_per_s || threads == 0 || threads > 256 || blocks == 0) return ZKB_ERR_INVALID_ARGUMENT;
    cudaSetDevice(ctx->device);
    ZKB_DISPATCH_ANY_FIELD(field, bench_field_mul, ctx, blocks, threads, iters, muls_per_s)
}

// Bare IMAD.WIDE issue-rate probe: shares no code with the Montgomery multiplier.  Every thread keeps 16 independent
// 64-bit accumulators and issues mad.wide.u32 on them round-robin (16 chains cover the dependent-issue latency), with a
// multiplicand that changes every round (one ALU add per 16 wide ops) so nothing folds.  wide multiply-adds/s measured
// here, divided by the wide count of a product (112 BLS12-381 Fr, 88 Pallas, 300 BLS12-381 Fq), is the ceiling the
// field-product peaks above are checked against.
__global__ void __launch_bounds__(256) bench_imad_wide_kernel(uint32_t iters, uint32_t seed, uint64_t *out) {
    uint64_t acc[16];
    uint32_t a[16];
    uint32_t b = seed + blockIdx.x * 40503u + 1u;
#pragma unroll
    for (int k = 0; k < 16; k++) {
        acc[k] = (uint64_t)(k + 1) * 0x9e3779b97f4a7c15ull;
        a[k] = (seed ^ (threadIdx.x * 2654435761u)) + 0x01000193u * (uint32_t)k;
    }
    for (uint32_t i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 2; r++) {
#pragma unroll
            for (int k = 0; k < 16; k++)
                asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(acc[k]) : "r"(a[k]), "r"(b));
            b += 0x9e3779b9u;
        }
    }
    uint64_t r = 0;
#pragma unroll
    for (int k = 0; k < 16; k++) r ^= acc[k];
    if (r == 0x123456789abcdef0ull) out[0] = r;   // practically never: keeps the chains live
}

extern "C" int zkb_bench_imad_wide(zkb_ctx *ctx, uint32_t blocks, uint32_t threads, uint32_t iters, double *wide_per_s) {
    if (!ctx || !wide_per_s || threads == 0 || threads > 256 || blocks == 0 || iters == 0) return ZKB_ERR_INVALID_ARGUMENT;
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    void *out;
    ZKB_TRY(ctx_scratch(ctx, "bench", 64, &out));
    cudaEvent_t e0, e1;
    ZKB_CUDA_OK(ctx, cudaEventCreate(&e0));
    ZKB_CUDA_OK(ctx, cudaEventCreate(&e1));
    bench_imad_wide_kernel<<<blocks, threads>>>(16, 12345u, (uint64_t *)out);   // warm-up
    ZKB_CUDA_OK(ctx, cudaEventRecord(e0));
    bench_imad_wide_kernel<<<blocks, threads>>>(iters, 12345u, (uint64_t *)out);
    ZKB_CUDA_OK(ctx, cudaEventRecord(e1));
    ZKB_CUDA_OK(ctx, cudaEventSynchronize(e1));
    ctx->launches += 2;
    float ms = 0;
    ZKB_CUDA_OK(ctx, cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *wide_per_s = 32.0 * (double)iters * blocks * threads / (ms * 1e-3);
    return ZKB_OK;
}

extern "C" void zkb_ctx_clear_error(zkb_ctx *ctx) {
    if (ctx) ctx->last_error.clear();
}

// ------------------------------------------------------------------------------------ device buffers for host templates
// The reference's commitment scheme keeps its polynomials as members between commit / eval_polys / proof_eval
// (zk/commitments/polynomial/lpc.hpp:66-200, batched_commitment.hpp:60-250); a host template over this ABI keeps them in
// device buffers instead.  Plain cudaMalloc / cudaMemcpyAsync / cudaMemsetAsync behind status codes, plus a gather of
// 32-byte elements (the FRI query phase reads 2 lambda values of every retained f_i, basic_fri.hpp:880-887).
__global__ void __launch_bounds__(256) gather_elems_kernel(const uint4 *__restrict__ src, uint32_t count,
                                                           const uint64_t *__restrict__ idx, uint4 *__restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * count) return;
    out[i] = src[2 * idx[i >> 1] + (i & 1)];
}

extern "C" {

int zkb_buf_alloc(zkb_ctx *ctx, uint64_t bytes, void **out) {
    if (!ctx || !out) return ZKB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaError_t e = cudaMalloc(out, bytes ? bytes : 1);
    if (e != cudaSuccess) {
        cudaGetLastError();
        *out = nullptr;
        return zkb::ctx_fail(ctx, ZKB_ERR_OUT_OF_MEMORY, std::string("zkb_buf_alloc: ") + cudaGetErrorString(e));
    }
    return ZKB_OK;
}

void zkb_buf_free(zkb_ctx *ctx, void *p) {
    if (!ctx || !p) return;
    cudaSetDevice(ctx->device);
    cudaFree(p);
}

int zkb_buf_copy(zkb_ctx *ctx, void *dst, int dst_mem, const void *src, int src_mem, uint64_t bytes, void *stream) {
    if (!ctx || (bytes && (!dst || !src))) return ZKB_ERR_INVALID_ARGUMENT;
    if (bytes == 0) return ZKB_OK;
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    cudaMemcpyKind kind = dst_mem == ZKB_MEM_DEVICE ? (src_mem == ZKB_MEM_DEVICE ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice)
                                                    : (src_mem == ZKB_MEM_DEVICE ? cudaMemcpyDeviceToHost : cudaMemcpyHostToHost);
    ZKB_CUDA_OK(ctx, cudaMemcpyAsync(dst, src, bytes, kind, st));
    if (kind != cudaMemcpyDeviceToDevice) ZKB_CUDA_OK(ctx, cudaStreamSynchronize(st));   // host buffers are the caller's again
    return ZKB_OK;
}

int zkb_buf_zero(zkb_ctx *ctx, void *dev, uint64_t bytes, void *stream) {
    if (!ctx || (bytes && !dev)) return ZKB_ERR_INVALID_ARGUMENT;
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    ZKB_CUDA_OK(ctx, cudaMemsetAsync(dev, 0, bytes, (cudaStream_t)stream));
    return ZKB_OK;
}

int zkb_gather(zkb_ctx *ctx, const void *src_device, uint32_t count, const uint64_t *indices, uint32_t *out, void *stream) {
    if (!ctx || (count && (!src_device || !indices || !out))) return ZKB_ERR_INVALID_ARGUMENT;
    if (count == 0) return ZKB_OK;
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    void *p;
    const size_t idx_pad = ((size_t)count * 8 + 15) & ~(size_t)15;
    ZKB_TRY(zkb::ctx_scratch(ctx, "gather", idx_pad + (size_t)count * 32, &p));
    uint64_t *d_idx = (uint64_t *)p;
    uint4 *d_out = (uint4 *)((char *)p + idx_pad);
    ZKB_CUDA_OK(ctx, cudaMemcpyAsync(d_idx, indices, (size_t)count * 8, cudaMemcpyHostToDevice, st));
    gather_elems_kernel<<<(2 * count + 255) / 256, 256, 0, st>>>((const uint4 *)src_device, count, d_idx, d_out);
    ctx->launches++;
    ZKB_CUDA_OK(ctx, cudaGetLastError());
    ZKB_CUDA_OK(ctx, cudaMemcpyAsync(out, d_out, (size_t)count * 32, cudaMemcpyDeviceToHost, st));
    ZKB_CUDA_OK(ctx, cudaStreamSynchronize(st));
    return ZKB_OK;
}

}  // extern "C"
