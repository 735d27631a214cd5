/* Oracle (TEST INFRASTRUCTURE ONLY): Jacobian short-Weierstrass (a = 0) group law + the BDLO12 bucket
 * MSM, restating crypto3-algebra's curve element `+`/`doubled()`/mixed add and
 * `multiexp<multiexp_method_BDLO12>` (libff lineage) as called from
 * zk/commitments/polynomial/kzg.hpp:146 and r1cs_gg_ppzksnark/prover.hpp:108-139. */
typedef struct { uint64_t X[NL], Y[NL], Z[NL]; } FN(jac);

static inline void FN(jzero)(const FN(field) * F, FN(jac) * r) { FN(fcopy)(r->X, F->r1); FN(fcopy)(r->Y, F->r1); FN(fzero)(r->Z); }
static inline int FN(jisinf)(const FN(jac) * p) { return FN(fiszero)(p->Z); }

static void FN(jdbl)(const FN(field) * F, FN(jac) * r, const FN(jac) * p) {
    if (FN(jisinf)(p)) { *r = *p; return; }
    uint64_t A[NL], B[NL], C[NL], D[NL], E[NL], G[NL], t[NL];
    FN(fmul)(F, A, p->X, p->X); FN(fmul)(F, B, p->Y, p->Y); FN(fmul)(F, C, B, B);
    FN(fadd)(F, t, p->X, B); FN(fmul)(F, t, t, t); FN(fsub)(F, t, t, A); FN(fsub)(F, t, t, C); FN(fadd)(F, D, t, t);
    FN(fadd)(F, E, A, A); FN(fadd)(F, E, E, A);
    FN(fmul)(F, G, E, E);
    uint64_t X3[NL], Y3[NL], Z3[NL];
    FN(fsub)(F, X3, G, D); FN(fsub)(F, X3, X3, D);
    FN(fadd)(F, C, C, C); FN(fadd)(F, C, C, C); FN(fadd)(F, C, C, C);
    FN(fsub)(F, t, D, X3); FN(fmul)(F, Y3, E, t); FN(fsub)(F, Y3, Y3, C);
    FN(fmul)(F, Z3, p->Y, p->Z); FN(fadd)(F, Z3, Z3, Z3);
    FN(fcopy)(r->X, X3); FN(fcopy)(r->Y, Y3); FN(fcopy)(r->Z, Z3);
}
static void FN(jadd)(const FN(field) * F, FN(jac) * r, const FN(jac) * p, const FN(jac) * q) {
    if (FN(jisinf)(p)) { *r = *q; return; }
    if (FN(jisinf)(q)) { *r = *p; return; }
    uint64_t Z1Z1[NL], Z2Z2[NL], U1[NL], U2[NL], S1[NL], S2[NL], H[NL], R[NL], HH[NL], HHH[NL], V[NL], t[NL];
    FN(fmul)(F, Z1Z1, p->Z, p->Z); FN(fmul)(F, Z2Z2, q->Z, q->Z);
    FN(fmul)(F, U1, p->X, Z2Z2); FN(fmul)(F, U2, q->X, Z1Z1);
    FN(fmul)(F, t, q->Z, Z2Z2); FN(fmul)(F, S1, p->Y, t);
    FN(fmul)(F, t, p->Z, Z1Z1); FN(fmul)(F, S2, q->Y, t);
    if (FN(feq)(U1, U2)) {
        if (FN(feq)(S1, S2)) { FN(jdbl)(F, r, p); return; }
        FN(jzero)(F, r); return;
    }
    FN(fsub)(F, H, U2, U1); FN(fsub)(F, R, S2, S1);
    FN(fmul)(F, HH, H, H); FN(fmul)(F, HHH, H, HH); FN(fmul)(F, V, U1, HH);
    uint64_t X3[NL], Y3[NL], Z3[NL];
    FN(fmul)(F, X3, R, R); FN(fsub)(F, X3, X3, HHH); FN(fsub)(F, X3, X3, V); FN(fsub)(F, X3, X3, V);
    FN(fsub)(F, t, V, X3); FN(fmul)(F, Y3, R, t); FN(fmul)(F, t, S1, HHH); FN(fsub)(F, Y3, Y3, t);
    FN(fmul)(F, Z3, p->Z, q->Z); FN(fmul)(F, Z3, Z3, H);
    FN(fcopy)(r->X, X3); FN(fcopy)(r->Y, Y3); FN(fcopy)(r->Z, Z3);
}
/* q affine (Z = 1), given as x,y in Montgomery form; all-zero (x,y) = infinity */
static void FN(jmadd)(const FN(field) * F, FN(jac) * r, const FN(jac) * p, const uint64_t *qx, const uint64_t *qy) {
    if (FN(fiszero)(qx) && FN(fiszero)(qy)) { *r = *p; return; }
    FN(jac) q;
    FN(fcopy)(q.X, qx); FN(fcopy)(q.Y, qy); FN(fcopy)(q.Z, F->r1);
    if (FN(jisinf)(p)) { *r = q; return; }
    uint64_t Z1Z1[NL], U2[NL], S2[NL], H[NL], R[NL], HH[NL], HHH[NL], V[NL], t[NL];
    FN(fmul)(F, Z1Z1, p->Z, p->Z); FN(fmul)(F, U2, qx, Z1Z1);
    FN(fmul)(F, t, p->Z, Z1Z1); FN(fmul)(F, S2, qy, t);
    if (FN(feq)(p->X, U2)) {
        if (FN(feq)(p->Y, S2)) { FN(jdbl)(F, r, p); return; }
        FN(jzero)(F, r); return;
    }
    FN(fsub)(F, H, U2, p->X); FN(fsub)(F, R, S2, p->Y);
    FN(fmul)(F, HH, H, H); FN(fmul)(F, HHH, H, HH); FN(fmul)(F, V, p->X, HH);
    uint64_t X3[NL], Y3[NL], Z3[NL];
    FN(fmul)(F, X3, R, R); FN(fsub)(F, X3, X3, HHH); FN(fsub)(F, X3, X3, V); FN(fsub)(F, X3, X3, V);
    FN(fsub)(F, t, V, X3); FN(fmul)(F, Y3, R, t); FN(fmul)(F, t, p->Y, HHH); FN(fsub)(F, Y3, Y3, t);
    FN(fmul)(F, Z3, p->Z, H);
    FN(fcopy)(r->X, X3); FN(fcopy)(r->Y, Y3); FN(fcopy)(r->Z, Z3);
}

/* One chunk of the bucket method: unsigned c-bit windows, most significant window first,
 * c doublings between windows, running-sum bucket reduction. points: Montgomery affine (x,y). */
static void FN(msm_chunk)(const FN(field) * F, const uint64_t *points, const uint64_t *scalars, size_t n, int scalar_bits,
                          FN(jac) * out) {
    FN(jac) result;
    FN(jzero)(F, &result);
    if (n == 0) { *out = result; return; }
    int log2n = 0;
    while (((size_t)2 << log2n) <= n) log2n++;
    int c = log2n - (log2n / 3 - 2);
    if (log2n < 6) c = log2n > 0 ? log2n : 1;
    if (c < 1) c = 1;
    int nwin = (scalar_bits + c - 1) / c;
    size_t nb = (size_t)1 << c;
    FN(jac) *buckets = (FN(jac) *)malloc(nb * sizeof(FN(jac)));
    for (int w = nwin - 1; w >= 0; w--) {
        for (int i = 0; i < c; i++) FN(jdbl)(F, &result, &result);
        for (size_t b = 0; b < nb; b++) FN(jzero)(F, &buckets[b]);
        int bit = w * c;
        for (size_t i = 0; i < n; i++) {
            const uint64_t *s = scalars + 4 * i;
            int limb = bit >> 6, sh = bit & 63;
            uint64_t d = s[limb] >> sh;
            if (sh + c > 64 && limb + 1 < 4) d |= s[limb + 1] << (64 - sh);
            d &= nb - 1;
            if (d) FN(jmadd)(F, &buckets[d], &buckets[d], points + 2 * NL * i, points + 2 * NL * i + NL);
        }
        FN(jac) running;
        FN(jzero)(F, &running);
        for (size_t b = nb - 1; b >= 1; b--) {
            FN(jadd)(F, &running, &running, &buckets[b]);
            FN(jadd)(F, &result, &result, &running);
        }
    }
    free(buckets);
    *out = result;
}

/* points_canon: n x (x,y) canonical; scalars: n x 4 limbs canonical; out: (x,y) canonical affine, zeros = inf.
 * chunks = threads, partial sums added (reference: `chunks` argument under MULTICORE, prover.hpp:94-99). */
static double FN(msm)(const FN(field) * F, const uint64_t *points_canon, const uint64_t *scalars, size_t n,
                      int scalar_bits, int threads, uint64_t *out_xy) {
    uint64_t *pm = (uint64_t *)malloc(n * 2 * NL * sizeof(uint64_t) + 8);
#pragma omp parallel for num_threads(threads) schedule(static)
    for (size_t i = 0; i < 2 * n; i++) FN(to_mont)(F, pm + NL * i, points_canon + NL * i);
    if (threads < 1) threads = 1;
    FN(jac) *partial = (FN(jac) *)malloc(threads * sizeof(FN(jac)));
    double t0 = now_s();
#pragma omp parallel for num_threads(threads) schedule(static, 1)
    for (int t = 0; t < threads; t++) {
        size_t lo = n * t / threads, hi = n * (t + 1) / threads;
        FN(msm_chunk)(F, pm + 2 * NL * lo, scalars + 4 * lo, hi - lo, scalar_bits, &partial[t]);
    }
    FN(jac) acc;
    FN(jzero)(F, &acc);
    for (int t = 0; t < threads; t++) FN(jadd)(F, &acc, &acc, &partial[t]);
    double dt = now_s() - t0;
    if (FN(jisinf)(&acc)) {
        for (int i = 0; i < 2 * NL; i++) out_xy[i] = 0;
    } else {
        uint64_t zi[NL], zi2[NL], zi3[NL], x[NL], y[NL];
        FN(finv)(F, zi, acc.Z); FN(fmul)(F, zi2, zi, zi); FN(fmul)(F, zi3, zi2, zi);
        FN(fmul)(F, x, acc.X, zi2); FN(fmul)(F, y, acc.Y, zi3);
        FN(from_mont)(F, out_xy, x); FN(from_mont)(F, out_xy + NL, y);
    }
    free(partial);
    free(pm);
    return dt;
}
