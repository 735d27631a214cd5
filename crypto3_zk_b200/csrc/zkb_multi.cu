// Multi-GPU entry points of the C ABI (SURVEY 8(b) "zkb_*_multi taking a device list", 8(e)): one host process drives
// several GPUs of a node - what a C++ prover linked against the reference's headers does; the one-process-per-GPU
// torch.distributed path (crypto3_zk_b200/sharding.py) partitions the same way.
//
//   MSM         contiguous point ranges per device, bases (and window tables) resident per device; one host thread per
//               device runs zkb_msm_partial on its slice of the scalars; the <= 192-byte XYZZ partials are added on the
//               host (zkb_msm_combine).  No device-to-device traffic.
//   LPC commit  polynomials split over the devices for the LDE; a Merkle leaf holds ALL polynomials at one index coset
//               (basic_fri.hpp:466-492), so the extended evaluations are regrouped by leaf range with peer copies over
//               NVLink - one strided cudaMemcpy2DAsync per (source, destination) pair, pulled by the destination -
//               every device hashes its leaf range and builds its subtree, and the top log2(devices) levels are hashed
//               from the subtree roots.  Same root as the single-GPU commit.
#include <string.h>
#include <thread>
#include <vector>
#include "zkb_internal.h"

using namespace zkb;

struct zkb_multi {
    std::vector<zkb_ctx *> ctx;
    std::string last_error;
};

struct zkb_msm_bases_multi {
    int curve = 0;
    uint64_t n = 0;
    std::vector<zkb_msm_bases *> part;     // part[d] holds points [off[d], off[d + 1])
    std::vector<uint64_t> off;
};

static int multi_fail(zkb_multi *m, int status, const std::string &msg) {
    if (m) m->last_error = msg;
    return status;
}

// runs fn(d) on one host thread per device and returns the first non-zero status
template <class Fn>
static int per_device(zkb_multi *m, Fn fn) {
    const size_t g = m->ctx.size();
    std::vector<int> st(g, ZKB_OK);
    std::vector<std::thread> th;
    for (size_t d = 1; d < g; d++) th.emplace_back([&, d] { st[d] = fn((uint32_t)d); });
    st[0] = fn(0);
    for (auto &t : th) t.join();
    for (size_t d = 0; d < g; d++)
        if (st[d] != ZKB_OK) {
            m->last_error = std::string("device ") + std::to_string(m->ctx[d]->device) + ": " + m->ctx[d]->last_error;
            return st[d];
        }
    return ZKB_OK;
}

extern "C" {

int zkb_multi_create(const int *devices, uint32_t count, zkb_multi **out) {
    if (!out) return ZKB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (!devices || count == 0) return ZKB_ERR_INVALID_ARGUMENT;
    for (uint32_t i = 0; i < count; i++)
        for (uint32_t j = 0; j < i; j++)
            if (devices[i] == devices[j]) return ZKB_ERR_INVALID_ARGUMENT;
    zkb_multi *m = new zkb_multi();
    for (uint32_t i = 0; i < count; i++) {
        zkb_ctx *c = nullptr;
        int s = zkb_ctx_create(devices[i], &c);
        if (s != ZKB_OK) {
            for (auto *x : m->ctx) zkb_ctx_destroy(x);
            delete m;
            return s;
        }
        m->ctx.push_back(c);
    }
    // peer access for the regroup copies (already-enabled / unsupported pairs fall back to staged copies)
    for (uint32_t i = 0; i < count; i++) {
        cudaSetDevice(devices[i]);
        for (uint32_t j = 0; j < count; j++) {
            if (i == j) continue;
            int can = 0;
            if (cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) == cudaSuccess && can) cudaDeviceEnablePeerAccess(devices[j], 0);
            cudaGetLastError();
        }
    }
    *out = m;
    return ZKB_OK;
}

void zkb_multi_destroy(zkb_multi *m) {
    if (!m) return;
    for (auto *c : m->ctx) zkb_ctx_destroy(c);
    delete m;
}

uint32_t zkb_multi_size(const zkb_multi *m) { return m ? (uint32_t)m->ctx.size() : 0; }
zkb_ctx *zkb_multi_ctx(zkb_multi *m, uint32_t i) { return m && i < m->ctx.size() ? m->ctx[i] : nullptr; }
const char *zkb_multi_last_error(const zkb_multi *m) { return m ? m->last_error.c_str() : ""; }

// ------------------------------------------------------------------------------------------------ MSM
int zkb_msm_bases_multi_create(zkb_multi *m, int curve, uint64_t n, const void *points_affine_host, zkb_msm_bases_multi **out) {
    if (!m || !out) return ZKB_ERR_INVALID_ARGUMENT;
    *out = nullptr;
    if (n && !points_affine_host) return multi_fail(m, ZKB_ERR_INVALID_ARGUMENT, "zkb_msm_bases_multi_create: null points");
    // bytes per affine point of the curve (x || y)
    uint32_t gen[48];
    if (zkb_curve_generator(curve, gen) != ZKB_OK) return multi_fail(m, ZKB_ERR_INVALID_ARGUMENT, "zkb_msm_bases_multi_create: bad curve");
    const size_t pb = curve == ZKB_CURVE_BLS12_381_G1 ? 96 : curve == ZKB_CURVE_BLS12_381_G2 ? 192 : curve == ZKB_CURVE_BN254_G2 ? 128 : 64;
    const uint32_t g = (uint32_t)m->ctx.size();
    zkb_msm_bases_multi *b = new zkb_msm_bases_multi();
    b->curve = curve;
    b->n = n;
    b->part.assign(g, nullptr);
    b->off.assign(g + 1, 0);
    for (uint32_t d = 0; d <= g; d++) b->off[d] = n / g * d + (d < n % g ? d : n % g);
    int s = per_device(m, [&](uint32_t d) {
        return zkb_msm_bases_create(m->ctx[d], curve, b->off[d + 1] - b->off[d], (const char *)points_affine_host + b->off[d] * pb,
                                    ZKB_MEM_HOST, nullptr, &b->part[d]);
    });
    if (s != ZKB_OK) {
        for (auto *p : b->part) zkb_msm_bases_free(p);
        delete b;
        return s;
    }
    *out = b;
    return ZKB_OK;
}

void zkb_msm_bases_multi_free(zkb_msm_bases_multi *b) {
    if (!b) return;
    for (auto *p : b->part) zkb_msm_bases_free(p);
    delete b;
}

int zkb_msm_bases_multi_precompute(zkb_multi *m, zkb_msm_bases_multi *b, int window_bits, uint64_t max_bytes_per_device) {
    if (!m || !b || b->part.size() != m->ctx.size()) return ZKB_ERR_INVALID_ARGUMENT;
    return per_device(m, [&](uint32_t d) { return zkb_msm_bases_precompute(m->ctx[d], b->part[d], window_bits, max_bytes_per_device, nullptr); });
}

int zkb_msm_multi(zkb_multi *m, const zkb_msm_bases_multi *b, uint64_t n, const void *scalars_host, uint32_t *result_affine) {
    if (!m || !b || !result_affine || b->part.size() != m->ctx.size()) return ZKB_ERR_INVALID_ARGUMENT;
    if (n > b->n || (n && !scalars_host)) return multi_fail(m, ZKB_ERR_INVALID_ARGUMENT, "zkb_msm_multi: more scalars than bases / null scalars");
    const uint32_t g = (uint32_t)m->ctx.size();
    std::vector<uint32_t> partials((size_t)g * 4 * 24, 0);
    uint32_t gen[48];
    zkb_curve_generator(b->curve, gen);
    const size_t words = b->curve == ZKB_CURVE_BLS12_381_G1 ? 48 : b->curve == ZKB_CURVE_BLS12_381_G2 ? 96 : b->curve == ZKB_CURVE_BN254_G2 ? 64 : 32;
    int s = per_device(m, [&](uint32_t d) {
        const uint64_t lo = b->off[d] < n ? b->off[d] : n, hi = b->off[d + 1] < n ? b->off[d + 1] : n;
        return zkb_msm_partial(m->ctx[d], b->part[d], 0, hi - lo, (const char *)scalars_host + lo * 32, ZKB_MEM_HOST,
                               partials.data() + (size_t)d * words, nullptr);
    });
    if (s != ZKB_OK) return s;
    return zkb_msm_combine(b->curve, g, partials.data(), result_affine);
}

// ------------------------------------------------------------------------------------------------ LPC commit
int zkb_lpc_commit_multi(zkb_multi *m, int field, int hash, int log_n_in, int log_n_out, int fri_step, uint32_t batch,
                         const void *polys_host, uint8_t *root_out) {
    if (!m || !polys_host || !root_out) return ZKB_ERR_INVALID_ARGUMENT;
    const uint32_t g = (uint32_t)m->ctx.size();
    if (g == 1) return zkb_lpc_commit(m->ctx[0], field, hash, log_n_in, log_n_out, fri_step, batch, polys_host, ZKB_MEM_HOST, root_out, nullptr, nullptr);
    int lg = 0;
    while ((1u << lg) < g) lg++;
    const int db = zkb_merkle_digest_bytes(hash);
    if ((1u << lg) != g || batch % g || batch == 0 || !db || log_n_in < 1 || log_n_out < log_n_in || fri_step < 1 || log_n_out - fri_step < lg)
        return multi_fail(m, ZKB_ERR_INVALID_ARGUMENT,
                          "zkb_lpc_commit_multi: the device count must be a power of two that divides the batch and the leaf count");
    const uint32_t pl = batch / g;                        // polynomials per device
    const uint64_t N = 1ull << log_n_out, L = N >> fri_step, lgc = L / g;   // leaves, leaves per device
    const uint64_t t = 1ull << fri_step;
    const size_t in_bytes = ((size_t)pl << log_n_in) * 32, ext_bytes = (size_t)pl * N * 32;
    std::vector<void *> ext(g, nullptr), recv(g, nullptr);
    // phase 1: upload and extend the local polynomials
    int s = per_device(m, [&](uint32_t d) {
        zkb_ctx *c = m->ctx[d];
        ZKB_CUDA_OK(c, cudaSetDevice(c->device));
        void *din;
        ZKB_TRY(ctx_scratch(c, "io_in", in_bytes, &din));
        ZKB_TRY(ctx_scratch(c, "multi_ext", ext_bytes, &ext[d]));
        ZKB_TRY(ctx_scratch(c, "multi_recv", (size_t)batch * (N / g) * 32, &recv[d]));
        ZKB_CUDA_OK(c, cudaMemcpyAsync(din, (const char *)polys_host + (size_t)d * in_bytes, in_bytes, cudaMemcpyHostToDevice, nullptr));
        ZKB_TRY(lde_device(c, field, log_n_in, log_n_out, pl, din, ext[d], nullptr));
        ZKB_CUDA_OK(c, cudaStreamSynchronize(nullptr));
        return (int)ZKB_OK;
    });
    if (s != ZKB_OK) return s;
    // phase 2: device d pulls, from every source, the rows (polynomial, coset half) restricted to its leaf range -
    // source row r = (p, tt) starts at element r L + d lgc (N = t L), destination rows are contiguous - then commits the
    // subtree of its leaves: the regrouped block is exactly [batch][N / g] evaluations with the same leaf pattern
    std::vector<uint8_t> roots((size_t)g * db);
    s = per_device(m, [&](uint32_t d) {
        zkb_ctx *c = m->ctx[d];
        ZKB_CUDA_OK(c, cudaSetDevice(c->device));
        for (uint32_t src = 0; src < g; src++) {
            const char *sp = (const char *)ext[src] + (size_t)d * lgc * 32;
            char *dp = (char *)recv[d] + (size_t)src * pl * t * lgc * 32;
            ZKB_CUDA_OK(c, cudaMemcpy2DAsync(dp, lgc * 32, sp, L * 32, lgc * 32, (size_t)pl * t, cudaMemcpyDefault, nullptr));
        }
        return merkle_build_device(c, hash, log_n_out - lg, fri_step, batch, recv[d], roots.data() + (size_t)d * db, nullptr, nullptr);
    });
    if (s != ZKB_OK) return s;
    return zkb_merkle_root_of_digests(m->ctx[0], hash, g, roots.data(), root_out, nullptr);
}

}  // extern "C"
