// Merkle commitment side of the LPC/FRI path: leaf packing + leaf hash fused in one kernel, then the
// binary tree.  Device replacement for the leaf loop and `containers::make_merkle_tree<Hash,2>` in
// zk/commitments/detail/polynomial/basic_fri.hpp:445-496 (precommit) and for lpc.hpp:101-106 (commit).
//
// Leaf x of a batch of L polynomials on a domain of size D with coset_size = 2^fri_step is the byte
// string  concat_{p < L} concat_{t < coset_size} BE32( poly_p[ s(x, t) ] ),
//   s(x, t) = x + bitrev(t >> 1, fri_step - 1) * (D >> fri_step) + (t & 1) * D/2
// which is the closed form of the s_indices recurrence at basic_fri.hpp:469-490 (checked against the
// oracle's literal restatement in tests).  Elements are canonical integers, 32 bytes big-endian
// (field_element_consumer.hpp:87-95, basic_fri.hpp:96-100).  Hashes: Keccak-256/512 with the
// original 0x01 padding (pinned by test/transcript/transcript.cpp:50-64) and SHA-256.
// The evaluations are read straight from the polynomial-major LDE output: consecutive threads take
// consecutive leaves, so every load instruction of a warp covers one contiguous 1 KiB run.
#include <stdio.h>
#include <string.h>
#include <vector>
#include "zkb_internal.h"

using namespace zkb;

struct zkb_merkle_tree {
    zkb_ctx *ctx;
    int device;        // the tree may be freed after its context: then the buffer goes straight back to the driver
    int hash;
    int digest_bytes;
    uint64_t leaves;
    uint8_t *d_nodes;  // level 0 (leaf digests) first, then each parent level; 2*leaves-1 digests
    size_t nodes_cap;  // capacity of d_nodes (it goes back to the context's tree pool)
};

// ------------------------------------------------------------------------------------ Keccak-f[1600]
__device__ __constant__ uint64_t KECCAK_RC[24] = {
    0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808Aull, 0x8000000080008000ull,
    0x000000000000808Bull, 0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull,
    0x000000000000008Aull, 0x0000000000000088ull, 0x0000000080008009ull, 0x000000008000000Aull,
    0x000000008000808Bull, 0x800000000000008Bull, 0x8000000000008089ull, 0x8000000000008003ull,
    0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800Aull, 0x800000008000000Aull,
    0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};

__device__ __forceinline__ uint64_t rol64(uint64_t x, int n) { return (x << n) | (x >> (64 - n)); }

// lanes a[x + 5 y]
__device__ __forceinline__ void keccak_f1600(uint64_t a[25]) {
#pragma unroll 1
    for (int r = 0; r < 24; r++) {
        uint64_t c0 = a[0] ^ a[5] ^ a[10] ^ a[15] ^ a[20];
        uint64_t c1 = a[1] ^ a[6] ^ a[11] ^ a[16] ^ a[21];
        uint64_t c2 = a[2] ^ a[7] ^ a[12] ^ a[17] ^ a[22];
        uint64_t c3 = a[3] ^ a[8] ^ a[13] ^ a[18] ^ a[23];
        uint64_t c4 = a[4] ^ a[9] ^ a[14] ^ a[19] ^ a[24];
        uint64_t d0 = c4 ^ rol64(c1, 1), d1 = c0 ^ rol64(c2, 1), d2 = c1 ^ rol64(c3, 1), d3 = c2 ^ rol64(c4, 1),
                 d4 = c3 ^ rol64(c0, 1);
        // theta + rho + pi: b[y + 5 ((2x+3y) % 5)] = rol(a[x + 5y] ^ d[x], ROT[x][y])
        uint64_t b0 = a[0] ^ d0;
        uint64_t b10 = rol64(a[1] ^ d1, 1);
        uint64_t b20 = rol64(a[2] ^ d2, 62);
        uint64_t b5 = rol64(a[3] ^ d3, 28);
        uint64_t b15 = rol64(a[4] ^ d4, 27);
        uint64_t b16 = rol64(a[5] ^ d0, 36);
        uint64_t b1 = rol64(a[6] ^ d1, 44);
        uint64_t b11 = rol64(a[7] ^ d2, 6);
        uint64_t b21 = rol64(a[8] ^ d3, 55);
        uint64_t b6 = rol64(a[9] ^ d4, 20);
        uint64_t b7 = rol64(a[10] ^ d0, 3);
        uint64_t b17 = rol64(a[11] ^ d1, 10);
        uint64_t b2 = rol64(a[12] ^ d2, 43);
        uint64_t b12 = rol64(a[13] ^ d3, 25);
        uint64_t b22 = rol64(a[14] ^ d4, 39);
        uint64_t b23 = rol64(a[15] ^ d0, 41);
        uint64_t b8 = rol64(a[16] ^ d1, 45);
        uint64_t b18 = rol64(a[17] ^ d2, 15);
        uint64_t b3 = rol64(a[18] ^ d3, 21);
        uint64_t b13 = rol64(a[19] ^ d4, 8);
        uint64_t b14 = rol64(a[20] ^ d0, 18);
        uint64_t b24 = rol64(a[21] ^ d1, 2);
        uint64_t b9 = rol64(a[22] ^ d2, 61);
        uint64_t b19 = rol64(a[23] ^ d3, 56);
        uint64_t b4 = rol64(a[24] ^ d4, 14);
        // chi
        a[0] = b0 ^ (~b1 & b2); a[1] = b1 ^ (~b2 & b3); a[2] = b2 ^ (~b3 & b4); a[3] = b3 ^ (~b4 & b0); a[4] = b4 ^ (~b0 & b1);
        a[5] = b5 ^ (~b6 & b7); a[6] = b6 ^ (~b7 & b8); a[7] = b7 ^ (~b8 & b9); a[8] = b8 ^ (~b9 & b5); a[9] = b9 ^ (~b5 & b6);
        a[10] = b10 ^ (~b11 & b12); a[11] = b11 ^ (~b12 & b13); a[12] = b12 ^ (~b13 & b14); a[13] = b13 ^ (~b14 & b10); a[14] = b14 ^ (~b10 & b11);
        a[15] = b15 ^ (~b16 & b17); a[16] = b16 ^ (~b17 & b18); a[17] = b17 ^ (~b18 & b19); a[18] = b18 ^ (~b19 & b15); a[19] = b19 ^ (~b15 & b16);
        a[20] = b20 ^ (~b21 & b22); a[21] = b21 ^ (~b22 & b23); a[22] = b22 ^ (~b23 & b24); a[23] = b23 ^ (~b24 & b20); a[24] = b24 ^ (~b20 & b21);
        a[0] ^= KECCAK_RC[r];
    }
}

__device__ __forceinline__ uint64_t bswap64(uint64_t v) {
    uint32_t lo = (uint32_t)v, hi = (uint32_t)(v >> 32);
    return ((uint64_t)__byte_perm(lo, 0, 0x0123) << 32) | __byte_perm(hi, 0, 0x0123);
}

// ------------------------------------------------------------------------------------ SHA-256
__device__ __constant__ uint32_t SHA256_K[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

__device__ __forceinline__ uint32_t ror32(uint32_t x, int n) { return __funnelshift_r(x, x, n); }

__device__ __forceinline__ void sha256_compress(uint32_t h[8], uint32_t w[16]) {
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        if (i >= 16) {
            uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
            uint32_t s0 = ror32(w15, 7) ^ ror32(w15, 18) ^ (w15 >> 3);
            uint32_t s1 = ror32(w2, 17) ^ ror32(w2, 19) ^ (w2 >> 10);
            w[i & 15] = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
        }
        uint32_t S1 = ror32(e, 6) ^ ror32(e, 11) ^ ror32(e, 25);
        uint32_t ch = (e & f) ^ (~e & g);
        uint32_t t1 = hh + S1 + ch + SHA256_K[i] + w[i & 15];
        uint32_t S0 = ror32(a, 2) ^ ror32(a, 13) ^ ror32(a, 22);
        uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t t2 = S0 + mj;
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}

__device__ __forceinline__ void sha256_init(uint32_t h[8]) {
    h[0] = 0x6a09e667; h[1] = 0xbb67ae85; h[2] = 0x3c6ef372; h[3] = 0xa54ff53a;
    h[4] = 0x510e527f; h[5] = 0x9b05688c; h[6] = 0x1f83d9ab; h[7] = 0x5be0cd19;
}

// ------------------------------------------------------------------------------------ leaf addressing
struct LeafGeom {
    const uint32_t *evals;   // [batch][D] elements of 8 limbs
    uint64_t D;              // domain size
    uint64_t leaves;         // D >> fri_step
    uint32_t batch;
    int fri_step;
    int log_d;
};

// pointer to the 8 limbs of message element e of leaf x
__device__ __forceinline__ const uint32_t *leaf_elem(const LeafGeom &g, uint64_t x, uint32_t e) {
    uint32_t cs = 1u << g.fri_step;
    uint32_t p = e >> g.fri_step, t = e & (cs - 1);
    uint32_t i = t >> 1;
    uint32_t rev = g.fri_step > 1 ? (__brev(i) >> (32 - (g.fri_step - 1))) : 0;
    uint64_t idx = x + (uint64_t)rev * (g.D >> g.fri_step) + (uint64_t)(t & 1) * (g.D >> 1);
    return g.evals + ((uint64_t)p * g.D + idx) * 8;
}

// Keccak lane m of the leaf message (8 message bytes, little-endian lane): element m/4, 64-bit word
// 3 - m%4 of the little-endian integer, byte-swapped (big-endian serialisation)
__device__ __forceinline__ uint64_t leaf_lane_keccak(const LeafGeom &g, uint64_t x, uint32_t m) {
    const uint32_t *el = leaf_elem(g, x, m >> 2);
    uint2 v = *reinterpret_cast<const uint2 *>(el + 2 * (3 - (m & 3)));
    return bswap64((uint64_t)v.x | ((uint64_t)v.y << 32));
}

template <int RATE_LANES, int DIGEST_LANES>
__global__ void __launch_bounds__(128) leaf_hash_keccak_kernel(LeafGeom g, uint64_t *__restrict__ digests) {
    uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= g.leaves) return;
    const uint32_t total_lanes = g.batch * (4u << g.fri_step);
    const uint32_t nblocks = total_lanes / RATE_LANES + 1;
    uint64_t a[25];
#pragma unroll
    for (int i = 0; i < 25; i++) a[i] = 0;
    for (uint32_t k = 0; k < nblocks; k++) {
        uint32_t m0 = k * RATE_LANES;
#pragma unroll
        for (int j = 0; j < RATE_LANES; j++) {
            uint32_t m = m0 + j;
            uint64_t lane = 0;
            if (m < total_lanes) lane = leaf_lane_keccak(g, x, m);
            else if (m == total_lanes) lane = 0x01ull;
            a[j] ^= lane;
        }
        if (k == nblocks - 1) a[RATE_LANES - 1] ^= 0x8000000000000000ull;
        keccak_f1600(a);
    }
#pragma unroll
    for (int i = 0; i < DIGEST_LANES; i++) digests[x * DIGEST_LANES + i] = a[i];
}

// parent[i] = Keccak(child[2i] || child[2i+1])
template <int RATE_LANES, int DIGEST_LANES>
__global__ void __launch_bounds__(128) node_hash_keccak_kernel(uint64_t parents, const uint64_t *__restrict__ child,
                                                               uint64_t *__restrict__ parent) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= parents) return;
    constexpr int MSG = 2 * DIGEST_LANES;
    constexpr int NBLK = MSG / RATE_LANES + 1;
    uint64_t a[25];
#pragma unroll
    for (int k = 0; k < 25; k++) a[k] = 0;
    const uint64_t *src = child + i * MSG;
#pragma unroll
    for (int k = 0; k < NBLK; k++) {
#pragma unroll
        for (int j = 0; j < RATE_LANES; j++) {
            int m = k * RATE_LANES + j;
            if (m < MSG) a[j] ^= src[m];
            else if (m == MSG) a[j] ^= 0x01ull;
        }
        if (k == NBLK - 1) a[RATE_LANES - 1] ^= 0x8000000000000000ull;
        keccak_f1600(a);
    }
#pragma unroll
    for (int k = 0; k < DIGEST_LANES; k++) parent[i * DIGEST_LANES + k] = a[k];
}

// SHA-256: 64-byte blocks = two elements; message word j of an element = limb[7 - j] (big-endian)
__global__ void __launch_bounds__(128) leaf_hash_sha256_kernel(LeafGeom g, uint32_t *__restrict__ digests) {
    uint64_t x = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= g.leaves) return;
    const uint32_t n_elems = g.batch << g.fri_step;   // always even (coset_size >= 2)
    uint32_t h[8], w[16];
    sha256_init(h);
    for (uint32_t e = 0; e < n_elems; e += 2) {
        const uint4 *p0 = reinterpret_cast<const uint4 *>(leaf_elem(g, x, e));
        const uint4 *p1 = reinterpret_cast<const uint4 *>(leaf_elem(g, x, e + 1));
        uint4 a0 = p0[0], a1 = p0[1], b0 = p1[0], b1 = p1[1];
        w[0] = a1.w; w[1] = a1.z; w[2] = a1.y; w[3] = a1.x; w[4] = a0.w; w[5] = a0.z; w[6] = a0.y; w[7] = a0.x;
        w[8] = b1.w; w[9] = b1.z; w[10] = b1.y; w[11] = b1.x; w[12] = b0.w; w[13] = b0.z; w[14] = b0.y; w[15] = b0.x;
        sha256_compress(h, w);
    }
    uint64_t bits = (uint64_t)n_elems * 256;
    w[0] = 0x80000000u;
#pragma unroll
    for (int i = 1; i < 14; i++) w[i] = 0;
    w[14] = (uint32_t)(bits >> 32);
    w[15] = (uint32_t)bits;
    sha256_compress(h, w);
    // digest bytes = big-endian words; store as little-endian uint32 so that memory order is the digest
#pragma unroll
    for (int i = 0; i < 8; i++) digests[x * 8 + i] = __byte_perm(h[i], 0, 0x0123);
}

__global__ void __launch_bounds__(128) node_hash_sha256_kernel(uint64_t parents, const uint32_t *__restrict__ child,
                                                               uint32_t *__restrict__ parent) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= parents) return;
    uint32_t h[8], w[16];
    sha256_init(h);
#pragma unroll
    for (int k = 0; k < 16; k++) w[k] = __byte_perm(child[i * 16 + k], 0, 0x0123);
    sha256_compress(h, w);
    w[0] = 0x80000000u;
#pragma unroll
    for (int k = 1; k < 15; k++) w[k] = 0;
    w[15] = 512;
    sha256_compress(h, w);
#pragma unroll
    for (int k = 0; k < 8; k++) parent[i * 8 + k] = __byte_perm(h[k], 0, 0x0123);
}

// ------------------------------------------------------------------------------------ grinding
// proof_of_work<TranscriptHash, uint32>::generate (zk/commitments/detail/polynomial/proof_of_work.hpp:47-68):
// find a nonce with (int_challenge(transcript + be32(nonce)) & mask) == 0, where for the sequential Fiat-Shamir
// transcript (zk/transcript/fiat_shamir.hpp:152-164,190-199) absorbing is state' = H(state || bytes) and
// int_challenge<uint32> is the low 32 bits of the big-endian integer H(state').  The reference walks the nonces
// one by one from a random start; here every thread tries one nonce and the smallest hit of the launch wins.
struct PowState {
    uint64_t lanes[8];   // transcript state: digest bytes as little-endian 64-bit lanes (keccak) / big-endian words packed (sha)
};

template <int RATE_LANES, int DIGEST_LANES>
__device__ __forceinline__ uint32_t pow_try_keccak(const PowState &st, uint32_t nonce) {
    uint64_t a[25];
#pragma unroll
    for (int i = 0; i < 25; i++) a[i] = 0;
#pragma unroll
    for (int i = 0; i < DIGEST_LANES; i++) a[i] = st.lanes[i];
    // the four nonce bytes (most significant first) follow the state; then the 0x01 pad byte
    a[DIGEST_LANES] = (uint64_t)__byte_perm(nonce, 0, 0x0123) | (0x01ull << 32);
    a[RATE_LANES - 1] ^= 0x8000000000000000ull;
    keccak_f1600(a);
#pragma unroll
    for (int i = DIGEST_LANES; i < 25; i++) a[i] = 0;
    a[DIGEST_LANES] = 0x01ull;
    a[RATE_LANES - 1] ^= 0x8000000000000000ull;
    keccak_f1600(a);
    // last four digest bytes, read as a big-endian integer
    return __byte_perm((uint32_t)(a[DIGEST_LANES - 1] >> 32), 0, 0x0123);
}

__device__ __forceinline__ uint32_t pow_try_sha256(const PowState &st, uint32_t nonce) {
    uint32_t h[8], w[16];
    sha256_init(h);
#pragma unroll
    for (int i = 0; i < 4; i++) { w[2 * i] = (uint32_t)(st.lanes[i] >> 32); w[2 * i + 1] = (uint32_t)st.lanes[i]; }
    w[8] = nonce;
    w[9] = 0x80000000u;
#pragma unroll
    for (int i = 10; i < 15; i++) w[i] = 0;
    w[15] = 36 * 8;
    sha256_compress(h, w);
#pragma unroll
    for (int i = 0; i < 8; i++) w[i] = h[i];
    w[8] = 0x80000000u;
#pragma unroll
    for (int i = 9; i < 15; i++) w[i] = 0;
    w[15] = 256;
    sha256_init(h);
    sha256_compress(h, w);
    return h[7];
}

template <int HASH>
__global__ void __launch_bounds__(256) pow_grind_kernel(PowState st, uint64_t base, uint64_t count, uint32_t mask,
                                                        unsigned long long *__restrict__ found) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    uint32_t nonce = (uint32_t)(base + i);
    uint32_t r;
    if (HASH == ZKB_HASH_KECCAK_256) r = pow_try_keccak<17, 4>(st, nonce);
    else if (HASH == ZKB_HASH_KECCAK_512) r = pow_try_keccak<9, 8>(st, nonce);
    else r = pow_try_sha256(st, nonce);
    if ((r & mask) == 0) atomicMin(found, (unsigned long long)(base + i));
}

// ------------------------------------------------------------------------------------ query phase: path gather
// out[q][d] = sibling digest of leaf indices[q] at tree level d (16-byte units; digests are 32 or 64 bytes)
__global__ void __launch_bounds__(256) merkle_paths_kernel(const uint4 *__restrict__ nodes, uint64_t leaves, int depth, int units,
                                                           uint32_t count, const uint64_t *__restrict__ indices,
                                                           uint4 *__restrict__ out) {
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t total = (uint64_t)count * depth * units;
    if (i >= total) return;
    uint32_t u = (uint32_t)(i % units);
    uint32_t d = (uint32_t)((i / units) % depth);
    uint32_t q = (uint32_t)(i / ((uint64_t)units * depth));
    // level d starts after leaves + leaves/2 + .. + leaves/2^(d-1) = 2 leaves (1 - 2^-d) digests
    uint64_t level_start = 2 * leaves - (2 * leaves >> d);
    uint64_t idx = (indices[q] >> d) ^ 1;
    out[i] = nodes[(level_start + idx) * units + u];
}

// ------------------------------------------------------------------------------------ host driver
static int digest_bytes_of(int hash) {
    switch (hash) {
        case ZKB_HASH_KECCAK_256: case ZKB_HASH_SHA2_256: return 32;
        case ZKB_HASH_KECCAK_512: return 64;
    }
    return 0;
}

static int merkle_build(zkb_ctx *ctx, int hash, int log_d, int fri_step, uint32_t batch, const void *d_evals,
                        uint8_t *root_out, zkb_merkle_tree **tree_out, cudaStream_t st) {
    return zkb::merkle_build_device(ctx, hash, log_d, fri_step, batch, d_evals, root_out, tree_out, st);
}
int zkb::merkle_build_device(zkb_ctx *ctx, int hash, int log_d, int fri_step, uint32_t batch, const void *d_evals,
                             uint8_t *root_out, zkb_merkle_tree **tree_out, cudaStream_t st) {
    const int db = digest_bytes_of(hash);
    const uint64_t D = 1ull << log_d, leaves = D >> fri_step;
    uint8_t *nodes = nullptr;
    size_t bytes = (size_t)(2 * leaves - 1) * db;
    bool keep = tree_out != nullptr;
    size_t nodes_cap = 0;
    if (keep) {
        void *p;
        ZKB_TRY(ctx_tree_alloc(ctx, bytes, &p, &nodes_cap));
        nodes = (uint8_t *)p;
    } else {
        void *p;
        ZKB_TRY(ctx_scratch(ctx, "merkle_nodes", bytes, &p));
        nodes = (uint8_t *)p;
    }
    LeafGeom g;
    g.evals = (const uint32_t *)d_evals;
    g.D = D;
    g.leaves = leaves;
    g.batch = batch;
    g.fri_step = fri_step;
    g.log_d = log_d;
    unsigned lb = (unsigned)((leaves + 127) / 128);
    switch (hash) {
        case ZKB_HASH_KECCAK_256: leaf_hash_keccak_kernel<17, 4><<<lb, 128, 0, st>>>(g, (uint64_t *)nodes); break;
        case ZKB_HASH_KECCAK_512: leaf_hash_keccak_kernel<9, 8><<<lb, 128, 0, st>>>(g, (uint64_t *)nodes); break;
        default: leaf_hash_sha256_kernel<<<lb, 128, 0, st>>>(g, (uint32_t *)nodes); break;
    }
    ctx->launches++;
    uint8_t *child = nodes;
    for (uint64_t n = leaves; n > 1; n >>= 1) {
        uint8_t *parent = child + n * db;
        uint64_t parents = n >> 1;
        unsigned pb = (unsigned)((parents + 127) / 128);
        switch (hash) {
            case ZKB_HASH_KECCAK_256: node_hash_keccak_kernel<17, 4><<<pb, 128, 0, st>>>(parents, (const uint64_t *)child, (uint64_t *)parent); break;
            case ZKB_HASH_KECCAK_512: node_hash_keccak_kernel<9, 8><<<pb, 128, 0, st>>>(parents, (const uint64_t *)child, (uint64_t *)parent); break;
            default: node_hash_sha256_kernel<<<pb, 128, 0, st>>>(parents, (const uint32_t *)child, (uint32_t *)parent); break;
        }
        ctx->launches++;
        child = parent;
    }
    cudaError_t ke = cudaGetLastError();
    if (ke == cudaSuccess && root_out) {
        ke = cudaMemcpyAsync(root_out, child, db, cudaMemcpyDeviceToHost, st);
        if (ke == cudaSuccess) ke = cudaStreamSynchronize(st);
    }
    if (ke != cudaSuccess) {
        if (keep) ctx_tree_release(ctx, nodes, nodes_cap);
        return ctx_fail(ctx, ZKB_ERR_CUDA, std::string("merkle build: ") + cudaGetErrorString(ke));
    }
    if (keep) {
        zkb_merkle_tree *t = new zkb_merkle_tree();
        t->ctx = ctx;
        t->device = ctx->device;
        t->hash = hash;
        t->digest_bytes = db;
        t->leaves = leaves;
        t->d_nodes = nodes;
        t->nodes_cap = nodes_cap;
        *tree_out = t;
    }
    return ZKB_OK;
}

extern "C" {

int zkb_merkle_digest_bytes(int hash) { return digest_bytes_of(hash); }
uint64_t zkb_merkle_leaves(const zkb_merkle_tree *t) { return t ? t->leaves : 0; }

void zkb_merkle_free(zkb_merkle_tree *t) {
    if (!t) return;
    if (t->d_nodes) {
        cudaSetDevice(t->device);
        if (zkb::ctx_alive(t->ctx)) zkb::ctx_tree_release(t->ctx, t->d_nodes, t->nodes_cap);
        else cudaFree(t->d_nodes);
    }
    delete t;
}

int zkb_merkle_path(zkb_ctx *ctx, const zkb_merkle_tree *t, uint64_t index, uint8_t *path_out) {
    if (!ctx || !t || !path_out || index >= t->leaves) return ZKB_ERR_INVALID_ARGUMENT;
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    const uint8_t *level = t->d_nodes;
    int d = 0;
    for (uint64_t n = t->leaves; n > 1; n >>= 1, d++) {
        ZKB_CUDA_OK(ctx, cudaMemcpy(path_out + (size_t)d * t->digest_bytes, level + (index ^ 1) * t->digest_bytes,
                                    t->digest_bytes, cudaMemcpyDeviceToHost));
        level += n * t->digest_bytes;
        index >>= 1;
    }
    return ZKB_OK;
}

int zkb_merkle_paths(zkb_ctx *ctx, const zkb_merkle_tree *t, uint32_t count, const uint64_t *indices, uint8_t *paths_out,
                     void *stream) {
    if (!ctx || !t || (count && (!indices || !paths_out))) return ZKB_ERR_INVALID_ARGUMENT;
    for (uint32_t q = 0; q < count; q++)
        if (indices[q] >= t->leaves) return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_merkle_paths: leaf index out of range");
    int depth = 0;
    for (uint64_t n = t->leaves; n > 1; n >>= 1) depth++;
    if (count == 0 || depth == 0) return ZKB_OK;
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int units = t->digest_bytes / 16;
    size_t out_bytes = (size_t)count * depth * t->digest_bytes, idx_bytes = (size_t)count * 8;
    void *p;
    const size_t idx_pad = (idx_bytes + 15) & ~(size_t)15;
    ZKB_TRY(ctx_scratch(ctx, "merkle_paths", idx_pad + out_bytes, &p));
    uint64_t *d_idx = (uint64_t *)p;
    uint4 *d_out = (uint4 *)((char *)p + idx_pad);
    ZKB_CUDA_OK(ctx, cudaMemcpyAsync(d_idx, indices, idx_bytes, cudaMemcpyHostToDevice, st));
    uint64_t total = (uint64_t)count * depth * units;
    merkle_paths_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>((const uint4 *)t->d_nodes, t->leaves, depth, units, count,
                                                                       d_idx, d_out);
    ctx->launches++;
    ZKB_CUDA_OK(ctx, cudaGetLastError());
    ZKB_CUDA_OK(ctx, cudaMemcpyAsync(paths_out, d_out, out_bytes, cudaMemcpyDeviceToHost, st));
    ZKB_CUDA_OK(ctx, cudaStreamSynchronize(st));
    return ZKB_OK;
}

int zkb_merkle_commit(zkb_ctx *ctx, int field, int hash, int log_n, int fri_step, uint32_t batch, const void *evals,
                      int mem, uint8_t *root_out, zkb_merkle_tree **tree_out, void *stream) {
    if (!ctx) return ZKB_ERR_INVALID_ARGUMENT;
    if (tree_out) *tree_out = nullptr;
    if (field < ZKB_FIELD_BLS12_381_FR || field > ZKB_FIELD_PALLAS_FQ || !digest_bytes_of(hash) || !evals || batch == 0 ||
        fri_step < 1 || log_n < fri_step || log_n > 40)
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_merkle_commit: need 1 <= fri_step <= log_n, batch >= 1");
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const void *d = evals;
    if (mem != ZKB_MEM_DEVICE) {
        void *p;
        size_t bytes = ((size_t)batch << log_n) * 32;
        ZKB_TRY(ctx_scratch(ctx, "io_out", bytes, &p));
        ZKB_CUDA_OK(ctx, cudaMemcpyAsync(p, evals, bytes, cudaMemcpyHostToDevice, st));
        d = p;
    }
    return merkle_build(ctx, hash, log_n, fri_step, batch, d, root_out, tree_out, st);
}

int zkb_merkle_root_of_digests(zkb_ctx *ctx, int hash, uint32_t count, const uint8_t *digests, uint8_t *root_out, void *stream) {
    if (!ctx) return ZKB_ERR_INVALID_ARGUMENT;
    const int db = digest_bytes_of(hash);
    if (!db || !digests || !root_out || count == 0 || (count & (count - 1)))
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_merkle_root_of_digests: count must be a power of two");
    if (count == 1) {
        memcpy(root_out, digests, db);
        return ZKB_OK;
    }
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    void *p;
    ZKB_TRY(ctx_scratch(ctx, "merkle_top", (size_t)2 * count * db, &p));
    uint8_t *child = (uint8_t *)p;
    ZKB_CUDA_OK(ctx, cudaMemcpyAsync(child, digests, (size_t)count * db, cudaMemcpyHostToDevice, st));
    for (uint64_t n = count; n > 1; n >>= 1) {
        uint8_t *parent = child + n * db;
        uint64_t parents = n >> 1;
        unsigned pb = (unsigned)((parents + 127) / 128);
        switch (hash) {
            case ZKB_HASH_KECCAK_256: node_hash_keccak_kernel<17, 4><<<pb, 128, 0, st>>>(parents, (const uint64_t *)child, (uint64_t *)parent); break;
            case ZKB_HASH_KECCAK_512: node_hash_keccak_kernel<9, 8><<<pb, 128, 0, st>>>(parents, (const uint64_t *)child, (uint64_t *)parent); break;
            default: node_hash_sha256_kernel<<<pb, 128, 0, st>>>(parents, (const uint32_t *)child, (uint32_t *)parent); break;
        }
        ctx->launches++;
        child = parent;
    }
    ZKB_CUDA_OK(ctx, cudaGetLastError());
    ZKB_CUDA_OK(ctx, cudaMemcpyAsync(root_out, child, db, cudaMemcpyDeviceToHost, st));
    ZKB_CUDA_OK(ctx, cudaStreamSynchronize(st));
    return ZKB_OK;
}

int zkb_lpc_commit(zkb_ctx *ctx, int field, int hash, int log_n_in, int log_n_out, int fri_step, uint32_t batch,
                   const void *polys, int mem, uint8_t *root_out, zkb_merkle_tree **tree_out, void *stream) {
    if (!ctx) return ZKB_ERR_INVALID_ARGUMENT;
    if (tree_out) *tree_out = nullptr;
    if (field < ZKB_FIELD_BLS12_381_FR || field > ZKB_FIELD_PALLAS_FQ || !digest_bytes_of(hash) || !polys || batch == 0 ||
        log_n_in < 1 || log_n_out < log_n_in || fri_step < 1 || log_n_out < fri_step)
        return ctx_fail(ctx, ZKB_ERR_INVALID_ARGUMENT, "zkb_lpc_commit: bad sizes");
    if (log_n_out > zkb_field_two_adicity(field)) return ctx_fail(ctx, ZKB_ERR_DOMAIN_TOO_LARGE, "2^log_n_out exceeds the two-adicity");
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    size_t ib = ((size_t)batch << log_n_in) * 32, ob = ((size_t)batch << log_n_out) * 32;
    // the extended evaluations are scratch: the reference does not retain them either
    // (precommit takes the container by value, basic_fri.hpp:445; lpc keeps only the tree, lpc.hpp:103)
    void *ext;
    ZKB_TRY(ctx_scratch(ctx, "lpc_ext", ob, &ext));
    if (mem == ZKB_MEM_DEVICE) {
        ZKB_TRY(lde_device(ctx, field, log_n_in, log_n_out, batch, polys, ext, st));
        return merkle_build(ctx, hash, log_n_out, fri_step, batch, ext, root_out, tree_out, st);
    }
    // host polynomials: uploaded in chunks on a copy stream, every chunk extended as soon as it has arrived (the upload of
    // config #2 is 2 GiB = 38 ms at PCIe rate, a quarter of the commit)
    void *p;
    ZKB_TRY(ctx_scratch(ctx, "io_in", ib, &p));
    if (!ctx->copy_in) {
        ZKB_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
        ZKB_CUDA_OK(ctx, cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; i++) {
            ZKB_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->ev_h2d[i], cudaEventDisableTiming));
            ZKB_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->ev_comp[i], cudaEventDisableTiming));
            ZKB_CUDA_OK(ctx, cudaEventCreateWithFlags(&ctx->ev_d2h[i], cudaEventDisableTiming));
        }
    }
    const size_t in_poly = (size_t)32 << log_n_in, out_poly = (size_t)32 << log_n_out;
    uint32_t chunk = (uint32_t)((256ull << 20) / in_poly);
    if (chunk < 1) chunk = 1;
    if (chunk > batch) chunk = batch;
    const uint32_t nchunks = (batch + chunk - 1) / chunk;
    std::vector<cudaEvent_t> ev(nchunks, nullptr);
    int status = ZKB_OK;
    cudaError_t e = cudaEventRecord(ctx->ev_comp[0], st);            // earlier work on st may still read io_in
    if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->copy_in, ctx->ev_comp[0], 0);
    for (uint32_t k = 0; k < nchunks && e == cudaSuccess; k++) {
        const uint32_t b0 = k * chunk, nb = batch - b0 < chunk ? batch - b0 : chunk;
        e = cudaEventCreateWithFlags(&ev[k], cudaEventDisableTiming);
        if (e == cudaSuccess)
            e = cudaMemcpyAsync((char *)p + b0 * in_poly, (const char *)polys + b0 * in_poly, nb * in_poly, cudaMemcpyHostToDevice, ctx->copy_in);
        if (e == cudaSuccess) e = cudaEventRecord(ev[k], ctx->copy_in);
    }
    for (uint32_t k = 0; k < nchunks && e == cudaSuccess && status == ZKB_OK; k++) {
        const uint32_t b0 = k * chunk, nb = batch - b0 < chunk ? batch - b0 : chunk;
        e = cudaStreamWaitEvent(st, ev[k], 0);
        if (e == cudaSuccess)
            status = lde_device(ctx, field, log_n_in, log_n_out, nb, (const char *)p + b0 * in_poly, (char *)ext + b0 * out_poly, st);
    }
    if (e == cudaSuccess && status == ZKB_OK) status = merkle_build(ctx, hash, log_n_out, fri_step, batch, ext, root_out, tree_out, st);
    cudaStreamSynchronize(ctx->copy_in);
    for (auto v : ev)
        if (v) cudaEventDestroy(v);
    if (e != cudaSuccess) return ctx_fail(ctx, ZKB_ERR_CUDA, std::string("zkb_lpc_commit: ") + cudaGetErrorString(e));
    return status;
}

int zkb_pow_grind(zkb_ctx *ctx, int hash, const uint8_t *state, uint32_t start, uint32_t mask, uint32_t *nonce_out,
                  void *stream) {
    if (!ctx || !state || !nonce_out || !digest_bytes_of(hash)) return ZKB_ERR_INVALID_ARGUMENT;
    ZKB_CUDA_OK(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int db = digest_bytes_of(hash);
    PowState ps;
    for (int i = 0; i < 8; i++) ps.lanes[i] = 0;
    for (int i = 0; i < db / 8; i++) {
        uint64_t v = 0;
        if (hash == ZKB_HASH_SHA2_256) for (int b = 0; b < 8; b++) v = (v << 8) | state[8 * i + b];           // two big-endian words
        else for (int b = 7; b >= 0; b--) v = (v << 8) | state[8 * i + b];                                      // little-endian lane
        ps.lanes[i] = v;
    }
    void *p;
    ZKB_TRY(ctx_scratch(ctx, "pow_found", 8, &p));
    unsigned long long *d_found = (unsigned long long *)p, h_found = ~0ull;
    ZKB_CUDA_OK(ctx, cudaMemcpyAsync(d_found, &h_found, 8, cudaMemcpyHostToDevice, st));
    // ascending chunks, so the first chunk with a hit holds the smallest nonce >= start
    uint64_t base = start, chunk = 1ull << 20;
    while (base <= 0xffffffffull) {
        uint64_t count = chunk;
        if (base + count > 0x100000000ull) count = 0x100000000ull - base;
        unsigned blocks = (unsigned)((count + 255) / 256);
        switch (hash) {
            case ZKB_HASH_KECCAK_256: pow_grind_kernel<ZKB_HASH_KECCAK_256><<<blocks, 256, 0, st>>>(ps, base, count, mask, d_found); break;
            case ZKB_HASH_KECCAK_512: pow_grind_kernel<ZKB_HASH_KECCAK_512><<<blocks, 256, 0, st>>>(ps, base, count, mask, d_found); break;
            default: pow_grind_kernel<ZKB_HASH_SHA2_256><<<blocks, 256, 0, st>>>(ps, base, count, mask, d_found); break;
        }
        ctx->launches++;
        ZKB_CUDA_OK(ctx, cudaGetLastError());
        ZKB_CUDA_OK(ctx, cudaMemcpyAsync(&h_found, d_found, 8, cudaMemcpyDeviceToHost, st));
        ZKB_CUDA_OK(ctx, cudaStreamSynchronize(st));
        if (h_found != ~0ull) {
            *nonce_out = (uint32_t)h_found;
            return ZKB_OK;
        }
        base += count;
        if (chunk < (1ull << 26)) chunk <<= 2;
    }
    return ctx_fail(ctx, ZKB_ERR_UNSUPPORTED, "zkb_pow_grind: no 32-bit nonce at or above `start` satisfies the mask");
}

}  // extern "C"
