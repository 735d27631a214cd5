// Internal (non-ABI) declarations shared by the translation units of libzkb200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <map>
#include <string>
#include <vector>
#include "../../include/zkb200.h"

struct zkb_ctx {
    int device = 0;
    int sm_count = 0;
    std::string last_error;
    uint64_t launches = 0;
    uint64_t scratch_limit = 6ull << 30;
    // grow-only scratch buffers, keyed by role
    struct Buf {
        void *p = nullptr;
        size_t cap = 0;
    };
    std::map<std::string, Buf> scratch;
    // cached device tables (twiddles, coset powers), keyed by a descriptive string
    std::map<std::string, Buf> tables;
    // buffers of freed Merkle trees, reused by the next kept tree: a prover builds and drops trees of the same few
    // sizes for every proof, and cudaMalloc / cudaFree cost milliseconds and synchronise the device
    std::vector<Buf> tree_pool;
    size_t tree_pool_bytes = 0;
    // copy streams + events of the host-buffer pipeline (created on first use, see host_pipeline())
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
};

namespace zkb {

int ctx_fail(zkb_ctx *ctx, int status, const std::string &msg);
// false once zkb_ctx_destroy has run on this pointer (handles that outlive their context must not touch it)
bool ctx_alive(const zkb_ctx *ctx);
// returns a device buffer of at least `bytes` (reallocates when too small)
int ctx_scratch(zkb_ctx *ctx, const char *role, size_t bytes, void **out);
// device memory of a kept Merkle tree: taken from / returned to ctx->tree_pool (bounded by the scratch limit)
int ctx_tree_alloc(zkb_ctx *ctx, size_t bytes, void **out, size_t *cap);
void ctx_tree_release(zkb_ctx *ctx, void *p, size_t cap);
// looks a table up; *created is set when the caller has to fill it
int ctx_table(zkb_ctx *ctx, const std::string &key, size_t bytes, void **out, bool *created);

#define ZKB_CUDA_OK(ctx, expr)                                                                       \
    do {                                                                                             \
        cudaError_t _e = (expr);                                                                     \
        if (_e != cudaSuccess)                                                                       \
            return zkb::ctx_fail(ctx, _e == cudaErrorMemoryAllocation ? ZKB_ERR_OUT_OF_MEMORY : ZKB_ERR_CUDA, \
                                 std::string(#expr) + ": " + cudaGetErrorString(_e));                \
    } while (0)

#define ZKB_TRY(expr)                 \
    do {                              \
        int _s = (expr);              \
        if (_s != ZKB_OK) return _s;  \
    } while (0)

// device-pointer level entry points used across translation units
int ntt_device(zkb_ctx *ctx, int field, int log_n, uint32_t batch, const void *d_in, void *d_out, int inverse,
               const uint32_t *coset_shift, uint64_t in_poly_stride, uint64_t in_valid_elems, cudaStream_t st,
               const void *known_src = nullptr, int known_log = 0, uint64_t known_poly_stride = 0);
// d_coef_out (optional, [batch][2^log_n_in]): receives the coefficient form the resize passes through
int lde_device(zkb_ctx *ctx, int field, int log_n_in, int log_n_out, uint32_t batch, const void *d_in, void *d_out,
               cudaStream_t st, void *d_coef_out = nullptr);
// one FRI fold on device buffers (alpha: host, canonical limbs)
int fold_device(zkb_ctx *ctx, int field, int log_n, const void *d_f, const uint32_t *alpha, void *d_out, cudaStream_t st);
// leaf packing + Merkle tree over extended evaluations [batch][2^log_d] on the device (zkb_hash.cu)
int merkle_build_device(zkb_ctx *ctx, int hash, int log_d, int fri_step, uint32_t batch, const void *d_evals,
                        uint8_t *root_out, zkb_merkle_tree **tree_out, cudaStream_t st);

}  // namespace zkb
