/* Oracle (TEST INFRASTRUCTURE ONLY): Montgomery field arithmetic on NL 64-bit limbs, included twice
 * (NL = 4 for the 254/255-bit fields, NL = 6 for BLS12-381 Fq) with FN(name) giving the suffix.
 * Restates what crypto3-multiprecision's modular_adaptor provides upstream:
 * values kept in Montgomery form, R = 2^(64 NL). */
typedef struct {
    uint64_t p[NL], r1[NL], r2[NL], ninv;
} FN(field);

static inline void FN(fadd)(const FN(field) * F, uint64_t *r, const uint64_t *a, const uint64_t *b) {
    unsigned __int128 c = 0;
    uint64_t t[NL];
    for (int i = 0; i < NL; i++) { c += (unsigned __int128)a[i] + b[i]; t[i] = (uint64_t)c; c >>= 64; }
    uint64_t s[NL];
    unsigned __int128 br = 0;
    for (int i = 0; i < NL; i++) { unsigned __int128 d = (unsigned __int128)t[i] - F->p[i] - (uint64_t)br; s[i] = (uint64_t)d; br = (d >> 64) & 1; }
    int ge = !br;
    for (int i = 0; i < NL; i++) r[i] = ge ? s[i] : t[i];
}
static inline void FN(fsub)(const FN(field) * F, uint64_t *r, const uint64_t *a, const uint64_t *b) {
    unsigned __int128 br = 0;
    uint64_t t[NL];
    for (int i = 0; i < NL; i++) { unsigned __int128 d = (unsigned __int128)a[i] - b[i] - (uint64_t)br; t[i] = (uint64_t)d; br = (d >> 64) & 1; }
    if (br) {
        unsigned __int128 c = 0;
        for (int i = 0; i < NL; i++) { c += (unsigned __int128)t[i] + F->p[i]; t[i] = (uint64_t)c; c >>= 64; }
    }
    for (int i = 0; i < NL; i++) r[i] = t[i];
}
static inline void FN(fmul)(const FN(field) * F, uint64_t *r, const uint64_t *a, const uint64_t *b) {
    uint64_t t[NL + 2];
    for (int i = 0; i < NL + 2; i++) t[i] = 0;
    for (int i = 0; i < NL; i++) {
        unsigned __int128 c = 0;
        for (int j = 0; j < NL; j++) { c += (unsigned __int128)a[j] * b[i] + t[j]; t[j] = (uint64_t)c; c >>= 64; }
        c += t[NL]; t[NL] = (uint64_t)c; t[NL + 1] = (uint64_t)(c >> 64);
        uint64_t m = t[0] * F->ninv;
        c = (unsigned __int128)m * F->p[0] + t[0]; c >>= 64;
        for (int j = 1; j < NL; j++) { c += (unsigned __int128)m * F->p[j] + t[j]; t[j - 1] = (uint64_t)c; c >>= 64; }
        c += t[NL]; t[NL - 1] = (uint64_t)c; t[NL] = t[NL + 1] + (uint64_t)(c >> 64);
    }
    uint64_t s[NL];
    unsigned __int128 br = 0;
    for (int i = 0; i < NL; i++) { unsigned __int128 d = (unsigned __int128)t[i] - F->p[i] - (uint64_t)br; s[i] = (uint64_t)d; br = (d >> 64) & 1; }
    int ge = t[NL] || !br;
    for (int i = 0; i < NL; i++) r[i] = ge ? s[i] : t[i];
}
static inline int FN(fiszero)(const uint64_t *a) { uint64_t x = 0; for (int i = 0; i < NL; i++) x |= a[i]; return x == 0; }
static inline int FN(feq)(const uint64_t *a, const uint64_t *b) { uint64_t x = 0; for (int i = 0; i < NL; i++) x |= a[i] ^ b[i]; return x == 0; }
static inline void FN(fcopy)(uint64_t *r, const uint64_t *a) { for (int i = 0; i < NL; i++) r[i] = a[i]; }
static inline void FN(fzero)(uint64_t *r) { for (int i = 0; i < NL; i++) r[i] = 0; }
static inline void FN(fneg)(const FN(field) * F, uint64_t *r, const uint64_t *a) {
    if (FN(fiszero)(a)) { FN(fzero)(r); return; }
    uint64_t z[NL]; FN(fzero)(z); FN(fsub)(F, r, z, a);
}
static inline void FN(to_mont)(const FN(field) * F, uint64_t *r, const uint64_t *a) { FN(fmul)(F, r, a, F->r2); }
static inline void FN(from_mont)(const FN(field) * F, uint64_t *r, const uint64_t *a) {
    uint64_t o[NL]; FN(fzero)(o); o[0] = 1; FN(fmul)(F, r, a, o);
}
static void FN(fpow)(const FN(field) * F, uint64_t *r, const uint64_t *a, const uint64_t *e, int elimbs) {
    uint64_t acc[NL], base[NL];
    FN(fcopy)(acc, F->r1); FN(fcopy)(base, a);
    for (int i = 0; i < elimbs; i++)
        for (int b = 0; b < 64; b++) {
            if ((e[i] >> b) & 1) FN(fmul)(F, acc, acc, base);
            FN(fmul)(F, base, base, base);
        }
    FN(fcopy)(r, acc);
}
static void FN(finv)(const FN(field) * F, uint64_t *r, const uint64_t *a) {
    uint64_t e[NL];
    for (int i = 0; i < NL; i++) e[i] = F->p[i];
    e[0] -= 2;
    FN(fpow)(F, r, a, e, NL);
}
/* builds the constants from the modulus */
static void FN(field_init)(FN(field) * F, const uint64_t *p) {
    for (int i = 0; i < NL; i++) F->p[i] = p[i];
    uint64_t x = 1;
    for (int i = 0; i < 6; i++) x *= 2 - p[0] * x;
    F->ninv = (uint64_t)0 - x;
    /* r1 = 2^(64 NL) mod p by repeated doubling of 1 */
    uint64_t t[NL]; FN(fzero)(t); t[0] = 1;
    for (int i = 0; i < 64 * NL; i++) FN(fadd)(F, t, t, t);
    FN(fcopy)(F->r1, t);
    for (int i = 0; i < 64 * NL; i++) FN(fadd)(F, t, t, t);
    FN(fcopy)(F->r2, t);
}
