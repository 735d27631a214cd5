/* CPU oracle in C - TEST INFRASTRUCTURE ONLY (used by tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference leg; never linked into or called from the product).
 *
 * Restates the algorithms crypto3-zk runs on the CPU for its two hot paths (the arithmetic itself is
 * in un-vendored sibling libraries, so this is a "port", not the reference binary):
 *   - basic_radix2_domain fft / inverse_fft: in-place bit-reversal + DIT butterflies with cached
 *     twiddles, inverse scaled by 1/m; multiply_by_coset (r1cs_to_qap.hpp:250-315)
 *   - polynomial_dfs::resize = iFFT, zero-pad, FFT (basic_fri.hpp:451-455)
 *   - precommit leaf packing + binary Merkle tree (basic_fri.hpp:445-496), keccak-256/512, sha-256
 *   - fold_polynomial dfs form (fold_polynomial.hpp:68-93)
 *   - multiexp<BDLO12> bucket method with `chunks` = threads (prover.hpp:94-99)
 * Parity pinning: see oracle/__init__.py.  This file is cross-checked against the Python oracle
 * (which is pinned to the reference's literal vectors) in tests/test_oracle_c.py.
 * Elements at this ABI: canonical little-endian 64-bit limbs (same bytes as the 32-bit limb ABI). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
}

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define NL 4
#define FN(name) CAT(name, 4)
#include "field_impl.h"
#include "curve_impl.h"
#undef NL
#undef FN
#define NL 6
#define FN(name) CAT(name, 6)
#include "field_impl.h"
#include "curve_impl.h"
#undef NL
#undef FN

/* ---- parameter table (same ids as include/zkb200.h) ---------------------------------------- */
typedef struct { int limbs; int bits; int two_adicity; uint64_t gen; uint64_t p[6]; } fparam;
static const fparam FIELDS[6] = {
    {4, 255, 32, 7, {0xffffffff00000001ull, 0x53bda402fffe5bfeull, 0x3339d80809a1d805ull, 0x73eda753299d7d48ull}},
    {4, 254, 28, 5, {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull}},
    {4, 255, 32, 5, {0x992d30ed00000001ull, 0x224698fc094cf91bull, 0x0000000000000000ull, 0x4000000000000000ull}},
    {4, 255, 32, 5, {0x8c46eb2100000001ull, 0x224698fc0994a8ddull, 0x0000000000000000ull, 0x4000000000000000ull}},
    {6, 381, 1, 2, {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull, 0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull}},
    {4, 254, 1, 3, {0x3c208c16d87cfd47ull, 0x97816a916871ca8dull, 0xb85045b68181585dull, 0x30644e72e131a029ull}},
};
/* curve id -> (base field id, scalar field id) */
static const int CURVE_BASE[3] = {4, 5, 2};
static const int CURVE_SCALAR[3] = {0, 1, 3};

int orc_threads_available(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ---- NTT --------------------------------------------------------------------------------------- */
static void omega4(const field4 *F, const fparam *fp, int log_n, uint64_t *w) {
    /* root_of_unity = g^((p-1)/2^s); unity_root(2^log_n) = root_of_unity^(2^(s-log_n)) */
    uint64_t g[4] = {fp->gen, 0, 0, 0}, e[4];
    to_mont4(F, g, g);
    /* e = (p - 1) >> s */
    uint64_t pm1[4];
    memcpy(pm1, F->p, sizeof pm1);
    pm1[0] -= 1;
    int s = fp->two_adicity;
    for (int i = 0; i < 4; i++) {
        int k = i + s / 64;
        uint64_t lo = k < 4 ? pm1[k] : 0, hi = k + 1 < 4 ? pm1[k + 1] : 0;
        e[i] = (s % 64) ? (lo >> (s % 64)) | (hi << (64 - s % 64)) : lo;
    }
    fpow4(F, w, g, e, 4);
    for (int i = 0; i < s - log_n; i++) fmul4(F, w, w, w);
}

/* in-place radix-2 DIT on Montgomery data, twiddles tw[j] = w^j (j < n/2) */
static void radix2_inplace(const field4 *F, uint64_t *a, int log_n, const uint64_t *tw) {
    size_t n = (size_t)1 << log_n;
    for (size_t k = 0; k < n; k++) {
        size_t r = 0;
        for (int b = 0; b < log_n; b++) r |= ((k >> b) & 1) << (log_n - 1 - b);
        if (k < r) {
            uint64_t t[4];
            memcpy(t, a + 4 * k, 32); memcpy(a + 4 * k, a + 4 * r, 32); memcpy(a + 4 * r, t, 32);
        }
    }
    for (int s = 1; s <= log_n; s++) {
        size_t m = (size_t)1 << (s - 1), stride = n >> s;
        for (size_t k = 0; k < n; k += 2 * m)
            for (size_t j = 0; j < m; j++) {
                uint64_t t[4], u[4];
                fmul4(F, t, a + 4 * (k + j + m), tw + 4 * (j * stride));
                memcpy(u, a + 4 * (k + j), 32);
                fadd4(F, a + 4 * (k + j), u, t);
                fsub4(F, a + 4 * (k + j + m), u, t);
            }
    }
}

static uint64_t *make_twiddles(const field4 *F, const uint64_t *w, int log_n) {
    size_t half = log_n ? (size_t)1 << (log_n - 1) : 1;
    uint64_t *tw = (uint64_t *)malloc(half * 32);
    memcpy(tw, F->r1, 32);
    for (size_t j = 1; j < half; j++) fmul4(F, tw + 4 * j, tw + 4 * (j - 1), w);
    return tw;
}

/* data: [batch][2^log_n] canonical, transformed in place.  shift (canonical) or NULL.
 * forward: a[i] *= g^i then fft; inverse: ifft (incl. 1/n) then a[i] *= g^-i.
 * Returns the seconds spent in the transforms proper (Montgomery conversion and twiddle set-up
 * excluded: upstream keeps Montgomery form and caches its twiddles). */
double orc_ntt(int field, int log_n, uint32_t batch, uint64_t *data, int inverse, const uint64_t *shift, int threads) {
    const fparam *fp = &FIELDS[field];
    field4 F;
    field_init4(&F, fp->p);
    size_t n = (size_t)1 << log_n;
    uint64_t w[4];
    omega4(&F, fp, log_n, w);
    if (inverse) finv4(&F, w, w);
    uint64_t *tw = make_twiddles(&F, w, log_n);
    uint64_t ninv[4] = {(uint64_t)n, 0, 0, 0}, g[4];
    to_mont4(&F, ninv, ninv);
    finv4(&F, ninv, ninv);
    if (shift) {
        to_mont4(&F, g, shift);
        if (inverse) finv4(&F, g, g);
    }
    if (threads < 1) threads = 1;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (size_t i = 0; i < (size_t)batch * n; i++) to_mont4(&F, data + 4 * i, data + 4 * i);
    double t0 = now_s();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (uint32_t b = 0; b < batch; b++) {
        uint64_t *a = data + 4 * n * b;
        if (shift && !inverse) {
            uint64_t u[4];
            memcpy(u, F.r1, 32);
            for (size_t i = 0; i < n; i++) { fmul4(&F, a + 4 * i, a + 4 * i, u); fmul4(&F, u, u, g); }
        }
        radix2_inplace(&F, a, log_n, tw);
        if (inverse) {
            for (size_t i = 0; i < n; i++) fmul4(&F, a + 4 * i, a + 4 * i, ninv);
            if (shift) {
                uint64_t u[4];
                memcpy(u, F.r1, 32);
                for (size_t i = 0; i < n; i++) { fmul4(&F, a + 4 * i, a + 4 * i, u); fmul4(&F, u, u, g); }
            }
        }
    }
    double dt = now_s() - t0;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (size_t i = 0; i < (size_t)batch * n; i++) from_mont4(&F, data + 4 * i, data + 4 * i);
    free(tw);
    return dt;
}

/* polynomial_dfs::resize for every polynomial: in [batch][2^log_in] -> out [batch][2^log_out] */
static double lde_mont(const field4 *F, const fparam *fp, int log_in, int log_out, uint32_t batch, const uint64_t *in_mont,
                       uint64_t *out_mont, int threads) {
    size_t nin = (size_t)1 << log_in, nout = (size_t)1 << log_out;
    uint64_t wi[4], wo[4];
    omega4(F, fp, log_in, wi);
    finv4(F, wi, wi);
    omega4(F, fp, log_out, wo);
    uint64_t *twi = make_twiddles(F, wi, log_in), *two = make_twiddles(F, wo, log_out);
    uint64_t ninv[4] = {(uint64_t)nin, 0, 0, 0};
    to_mont4(F, ninv, ninv);
    finv4(F, ninv, ninv);
    double t0 = now_s();
#pragma omp parallel for num_threads(threads) schedule(dynamic, 1)
    for (uint32_t b = 0; b < batch; b++) {
        uint64_t *a = out_mont + 4 * nout * b;
        memcpy(a, in_mont + 4 * nin * b, nin * 32);
        if (log_in != log_out) {
            radix2_inplace(F, a, log_in, twi);
            for (size_t i = 0; i < nin; i++) fmul4(F, a + 4 * i, a + 4 * i, ninv);
            memset(a + 4 * nin, 0, (nout - nin) * 32);
            radix2_inplace(F, a, log_out, two);
        }
    }
    double dt = now_s() - t0;
    free(twi);
    free(two);
    return dt;
}

double orc_lde(int field, int log_in, int log_out, uint32_t batch, const uint64_t *in, uint64_t *out, int threads) {
    const fparam *fp = &FIELDS[field];
    field4 F;
    field_init4(&F, fp->p);
    size_t nin = (size_t)1 << log_in, nout = (size_t)1 << log_out;
    if (threads < 1) threads = 1;
    uint64_t *im = (uint64_t *)malloc((size_t)batch * nin * 32);
#pragma omp parallel for num_threads(threads) schedule(static)
    for (size_t i = 0; i < (size_t)batch * nin; i++) to_mont4(&F, im + 4 * i, in + 4 * i);
    double dt = lde_mont(&F, fp, log_in, log_out, batch, im, out, threads);
#pragma omp parallel for num_threads(threads) schedule(static)
    for (size_t i = 0; i < (size_t)batch * nout; i++) from_mont4(&F, out + 4 * i, out + 4 * i);
    free(im);
    return dt;
}

/* fold_polynomial (dfs): out[i] = 1/2((1 + a w^-i) f[i] + (1 - a w^-i) f[i + n/2]), serial acc *= w^-1 */
double orc_fri_fold(int field, int log_n, const uint64_t *f, const uint64_t *alpha, uint64_t *out) {
    const fparam *fp = &FIELDS[field];
    field4 F;
    field_init4(&F, fp->p);
    size_t n = (size_t)1 << log_n, half = n / 2;
    uint64_t winv[4], acc[4], two_inv[4] = {2, 0, 0, 0};
    omega4(&F, fp, log_n, winv);
    finv4(&F, winv, winv);
    to_mont4(&F, two_inv, two_inv);
    finv4(&F, two_inv, two_inv);
    to_mont4(&F, acc, alpha);
    uint64_t *fm = (uint64_t *)malloc(n * 32);
    for (size_t i = 0; i < n; i++) to_mont4(&F, fm + 4 * i, f + 4 * i);
    double t0 = now_s();
    for (size_t i = 0; i < half; i++) {
        uint64_t p1[4], m1[4], t[4], u[4];
        fadd4(&F, p1, F.r1, acc);
        fsub4(&F, m1, F.r1, acc);
        fmul4(&F, t, p1, fm + 4 * i);
        fmul4(&F, u, m1, fm + 4 * (i + half));
        fadd4(&F, t, t, u);
        fmul4(&F, out + 4 * i, t, two_inv);
        fmul4(&F, acc, acc, winv);
    }
    double dt = now_s() - t0;
    for (size_t i = 0; i < half; i++) from_mont4(&F, out + 4 * i, out + 4 * i);
    free(fm);
    return dt;
}

/* ---- hashes ------------------------------------------------------------------------------------- */
static const uint64_t KRC[24] = {
    0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808Aull, 0x8000000080008000ull, 0x000000000000808Bull,
    0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008Aull, 0x0000000000000088ull,
    0x0000000080008009ull, 0x000000008000000Aull, 0x000000008000808Bull, 0x800000000000008Bull, 0x8000000000008089ull,
    0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800Aull, 0x800000008000000Aull,
    0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
static const int KROT[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};

static void keccak_f(uint64_t a[25]) {
    for (int r = 0; r < 24; r++) {
        uint64_t c[5], d[5], b[25];
        for (int x = 0; x < 5; x++) c[x] = a[x] ^ a[x + 5] ^ a[x + 10] ^ a[x + 15] ^ a[x + 20];
        for (int x = 0; x < 5; x++) {
            uint64_t t = c[(x + 1) % 5];
            d[x] = c[(x + 4) % 5] ^ ((t << 1) | (t >> 63));
        }
        for (int y = 0; y < 5; y++)
            for (int x = 0; x < 5; x++) {
                uint64_t v = a[x + 5 * y] ^ d[x];
                int rot = KROT[x + 5 * y];
                if (rot) v = (v << rot) | (v >> (64 - rot));
                b[y + 5 * ((2 * x + 3 * y) % 5)] = v;
            }
        for (int y = 0; y < 5; y++)
            for (int x = 0; x < 5; x++) a[x + 5 * y] = b[x + 5 * y] ^ (~b[(x + 1) % 5 + 5 * y] & b[(x + 2) % 5 + 5 * y]);
        a[0] ^= KRC[r];
    }
}
/* original Keccak padding 0x01 (pinned by test/transcript/transcript.cpp:50-64) */
static void keccak(const uint8_t *msg, size_t len, int rate, uint8_t *out, int outlen) {
    uint64_t a[25];
    memset(a, 0, sizeof a);
    uint8_t blk[144];
    while (1) {
        size_t take = len < (size_t)rate ? len : (size_t)rate;
        int last = len < (size_t)rate;
        memset(blk, 0, sizeof blk);
        memcpy(blk, msg, take);
        if (last) { blk[take] ^= 0x01; blk[rate - 1] ^= 0x80; }
        for (int i = 0; i < rate / 8; i++) {
            uint64_t v;
            memcpy(&v, blk + 8 * i, 8);
            a[i] ^= v;
        }
        keccak_f(a);
        msg += take; len -= take;
        if (last) break;
    }
    memcpy(out, a, outlen);
}
static const uint32_t SK[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
#define ROR(x, n) (((x) >> (n)) | ((x) << (32 - (n))))
static void sha256_block(uint32_t h[8], const uint8_t *blk) {
    uint32_t w[64];
    for (int i = 0; i < 16; i++) w[i] = ((uint32_t)blk[4 * i] << 24) | ((uint32_t)blk[4 * i + 1] << 16) | ((uint32_t)blk[4 * i + 2] << 8) | blk[4 * i + 3];
    for (int i = 16; i < 64; i++) {
        uint32_t s0 = ROR(w[i - 15], 7) ^ ROR(w[i - 15], 18) ^ (w[i - 15] >> 3), s1 = ROR(w[i - 2], 17) ^ ROR(w[i - 2], 19) ^ (w[i - 2] >> 10);
        w[i] = w[i - 16] + s0 + w[i - 7] + s1;
    }
    uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
    for (int i = 0; i < 64; i++) {
        uint32_t t1 = hh + (ROR(e, 6) ^ ROR(e, 11) ^ ROR(e, 25)) + ((e & f) ^ (~e & g)) + SK[i] + w[i];
        uint32_t t2 = (ROR(a, 2) ^ ROR(a, 13) ^ ROR(a, 22)) + ((a & b) ^ (a & c) ^ (b & c));
        hh = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
    }
    h[0] += a; h[1] += b; h[2] += c; h[3] += d; h[4] += e; h[5] += f; h[6] += g; h[7] += hh;
}
static void sha256(const uint8_t *msg, size_t len, uint8_t *out) {
    uint32_t h[8] = {0x6a09e667, 0xbb67ae85, 0x3c6ef372, 0xa54ff53a, 0x510e527f, 0x9b05688c, 0x1f83d9ab, 0x5be0cd19};
    size_t full = len / 64;
    for (size_t i = 0; i < full; i++) sha256_block(h, msg + 64 * i);
    uint8_t tail[128];
    size_t rem = len - 64 * full;
    memset(tail, 0, sizeof tail);
    memcpy(tail, msg + 64 * full, rem);
    tail[rem] = 0x80;
    size_t tl = rem + 9 <= 64 ? 64 : 128;
    uint64_t bits = (uint64_t)len * 8;
    for (int i = 0; i < 8; i++) tail[tl - 1 - i] = (uint8_t)(bits >> (8 * i));
    sha256_block(h, tail);
    if (tl == 128) sha256_block(h, tail + 64);
    for (int i = 0; i < 8; i++) { out[4 * i] = h[i] >> 24; out[4 * i + 1] = h[i] >> 16; out[4 * i + 2] = h[i] >> 8; out[4 * i + 3] = h[i]; }
}
static int digest_len(int hash) { return hash == 2 ? 64 : 32; }
static void hash_bytes(int hash, const uint8_t *msg, size_t len, uint8_t *out) {
    if (hash == 0) keccak(msg, len, 136, out, 32);
    else if (hash == 1) sha256(msg, len, out);
    else keccak(msg, len, 72, out, 64);
}
void orc_hash(int hash, const uint8_t *msg, size_t len, uint8_t *out) { hash_bytes(hash, msg, len, out); }

/* ---- precommit: leaves (basic_fri.hpp:466-492) + make_merkle_tree ------------------------------- */
/* evals: [batch][D] canonical.  nodes_out (optional): all levels, leaves first ((2*leaves-1)*digest). */
double orc_merkle_commit(int hash, int log_d, int fri_step, uint32_t batch, const uint64_t *evals, uint8_t *root_out,
                         uint8_t *nodes_out, int threads) {
    size_t D = (size_t)1 << log_d, coset = (size_t)1 << fri_step, leaves = D / coset;
    int dl = digest_len(hash);
    uint8_t *nodes = nodes_out ? nodes_out : (uint8_t *)malloc((2 * leaves - 1) * dl);
    size_t leaf_len = (size_t)batch * coset * 32;
    if (threads < 1) threads = 1;
    double t0 = now_s();
#pragma omp parallel num_threads(threads)
    {
        uint8_t *buf = (uint8_t *)malloc(leaf_len);
        size_t *s_idx = (size_t *)malloc(coset * sizeof(size_t));
#pragma omp for schedule(static)
        for (size_t x = 0; x < leaves; x++) {
            uint8_t *cur = buf;
            for (uint32_t p = 0; p < batch; p++) {
                /* the reference's s_indices recurrence, literally */
                s_idx[0] = x;
                s_idx[1] = (x + D / 2) % D;
                size_t base_index = D / 4, prev_half = 1, i = 1;
                while (i < coset / 2) {
                    for (size_t j = 0; j < prev_half; j++) {
                        s_idx[2 * i] = (base_index + s_idx[2 * j]) % D;
                        s_idx[2 * i + 1] = (s_idx[2 * i] + D / 2) % D;
                        i++;
                    }
                    base_index /= 2;
                    prev_half <<= 1;
                }
                for (size_t t = 0; t < coset; t++) {
                    const uint64_t *e = evals + 4 * ((size_t)p * D + s_idx[t]);
                    for (int k = 0; k < 4; k++)
                        for (int bb = 0; bb < 8; bb++) cur[8 * k + bb] = (uint8_t)(e[3 - k] >> (56 - 8 * bb));
                    cur += 32;
                }
            }
            hash_bytes(hash, buf, leaf_len, nodes + x * dl);
        }
        free(buf);
        free(s_idx);
    }
    uint8_t *child = nodes;
    for (size_t n = leaves; n > 1; n >>= 1) {
        uint8_t *parent = child + n * dl;
#pragma omp parallel for num_threads(threads) schedule(static)
        for (size_t i = 0; i < n / 2; i++) hash_bytes(hash, child + 2 * i * dl, 2 * dl, parent + i * dl);
        child = parent;
    }
    double dt = now_s() - t0;
    memcpy(root_out, child, dl);
    if (!nodes_out) free(nodes);
    return dt;
}

/* lpc commit = resize every polynomial to 2^log_out, then the tree.  Returns total seconds;
 * *lde_seconds (optional) receives the LDE share. */
double orc_lpc_commit(int field, int hash, int log_in, int log_out, int fri_step, uint32_t batch, const uint64_t *polys,
                      uint8_t *root_out, int threads, double *lde_seconds) {
    size_t nout = (size_t)1 << log_out;
    uint64_t *ext = (uint64_t *)malloc((size_t)batch * nout * 32);
    double t1 = orc_lde(field, log_in, log_out, batch, polys, ext, threads);
    double t2 = orc_merkle_commit(hash, log_out, fri_step, batch, ext, root_out, NULL, threads);
    free(ext);
    if (lde_seconds) *lde_seconds = t1;
    return t1 + t2;
}

/* ---- MSM ------------------------------------------------------------------------------------------ */
double orc_msm(int curve, size_t n, const uint64_t *points, const uint64_t *scalars, uint64_t *out_xy, int threads) {
    const fparam *bf = &FIELDS[CURVE_BASE[curve]];
    int sbits = FIELDS[CURVE_SCALAR[curve]].bits;
    if (threads < 1) threads = 1;
    if ((size_t)threads > n && n > 0) threads = (int)n;
    if (bf->limbs == 6) {
        field6 F;
        field_init6(&F, bf->p);
        return msm6(&F, points, scalars, n, sbits, threads, out_xy);
    }
    field4 F;
    field_init4(&F, bf->p);
    return msm4(&F, points, scalars, n, sbits, threads, out_xy);
}
