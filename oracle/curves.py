"""Short-Weierstrass groups + MSM restatement (oracle; test infrastructure only).

crypto3-algebra's curves / `algebra::multiexp` are un-vendored; this follows the published
algorithms and is pinned by the literal BLS12-381 vectors in
test/systems/ppzksnark/r1cs_gg_ppzksnark/r1cs_gg_ppzksnark_aggregation_conformity.cpp
(:864-930 prove_commitment, :1065-1884 gipa) - see tests/test_oracle_golden.py.
Call sites restated: kzg.hpp:100-118,143-148 (SRS powers + commit), ipp2/srs.hpp:42-55,
knowledge_commitment_multiexp.hpp:57-108 (0/1 pre-filter), r1cs_gg_ppzksnark/prover.hpp:108-139.

Group results are algorithm independent: equality is checked in affine form.
Points: None = infinity, else (x, y) affine with coordinates in the base field
(int for Fq, (c0, c1) tuples for Fq2).
"""
from .fields import BLS12_381_FQ, BLS12_381_FR, BN254_FQ, BN254_FR, PALLAS_FP, PALLAS_FQ


# --------------------------------------------------------------------------- base-field op tables
class FqOps:
    def __init__(self, p):
        self.p = p
        self.zero, self.one = 0, 1

    def add(self, a, b): return (a + b) % self.p
    def sub(self, a, b): return (a - b) % self.p
    def mul(self, a, b): return a * b % self.p
    def neg(self, a): return (-a) % self.p
    def inv(self, a): return pow(a, self.p - 2, self.p)
    def small(self, k): return k % self.p
    def is_zero(self, a): return a % self.p == 0


class Fq2Ops:
    """Fq[u]/(u^2 - nr) with nr = -1 for both BLS12-381 and BN254."""

    def __init__(self, p, nr=-1):
        self.p, self.nr = p, nr % p
        self.zero, self.one = (0, 0), (1, 0)

    def add(self, a, b): return ((a[0] + b[0]) % self.p, (a[1] + b[1]) % self.p)
    def sub(self, a, b): return ((a[0] - b[0]) % self.p, (a[1] - b[1]) % self.p)
    def neg(self, a): return ((-a[0]) % self.p, (-a[1]) % self.p)

    def mul(self, a, b):
        p = self.p
        return ((a[0] * b[0] + self.nr * a[1] * b[1]) % p, (a[0] * b[1] + a[1] * b[0]) % p)

    def inv(self, a):
        p = self.p
        d = pow((a[0] * a[0] - self.nr * a[1] * a[1]) % p, p - 2, p)
        return (a[0] * d % p, (-a[1]) * d % p)

    def small(self, k): return (k % self.p, 0)
    def is_zero(self, a): return a[0] % self.p == 0 and a[1] % self.p == 0


class Curve:
    """y^2 = x^3 + b (a = 0 for every curve on the hot path)."""

    def __init__(self, name, cid, F, b, gen, scalar_field, base_field, coord_limbs32):
        self.name, self.cid, self.F, self.b, self.gen = name, cid, F, b, gen
        self.scalar_field, self.base_field = scalar_field, base_field
        self.coord_limbs32 = coord_limbs32        # u32 limbs per coordinate (x or y) at the C ABI

    # ---- affine
    def is_on_curve(self, P):
        if P is None:
            return True
        F = self.F
        x, y = P
        return F.sub(F.mul(y, y), F.add(F.mul(F.mul(x, x), x), self.b)) == F.zero

    def neg(self, P):
        return None if P is None else (P[0], self.F.neg(P[1]))

    def add(self, P, Q):
        F = self.F
        if P is None:
            return Q
        if Q is None:
            return P
        if P[0] == Q[0]:
            if F.is_zero(F.add(P[1], Q[1])):
                return None
            lam = F.mul(F.mul(F.small(3), F.mul(P[0], P[0])), F.inv(F.mul(F.small(2), P[1])))
        else:
            lam = F.mul(F.sub(Q[1], P[1]), F.inv(F.sub(Q[0], P[0])))
        x3 = F.sub(F.sub(F.mul(lam, lam), P[0]), Q[0])
        y3 = F.sub(F.mul(lam, F.sub(P[0], x3)), P[1])
        return (x3, y3)

    # ---- Jacobian (X, Y, Z), a = 0; used for speed in python and as the restated formulas
    def j_from_affine(self, P):
        F = self.F
        return (F.one, F.one, F.zero) if P is None else (P[0], P[1], F.one)

    def j_to_affine(self, J):
        F = self.F
        X, Y, Z = J
        if F.is_zero(Z):
            return None
        zi = F.inv(Z)
        zi2 = F.mul(zi, zi)
        return (F.mul(X, zi2), F.mul(Y, F.mul(zi2, zi)))

    def j_double(self, J):
        F = self.F
        X, Y, Z = J
        if F.is_zero(Z):
            return J
        A = F.mul(X, X)
        B = F.mul(Y, Y)
        C = F.mul(B, B)
        t = F.add(X, B)
        D = F.sub(F.sub(F.mul(t, t), A), C)
        D = F.add(D, D)
        E = F.add(F.add(A, A), A)
        Fv = F.mul(E, E)
        X3 = F.sub(Fv, F.add(D, D))
        C8 = F.mul(F.small(8), C)
        Y3 = F.sub(F.mul(E, F.sub(D, X3)), C8)
        Z3 = F.mul(F.add(Y, Y), Z)
        return (X3, Y3, Z3)

    def j_add(self, P, Q):
        F = self.F
        if F.is_zero(P[2]):
            return Q
        if F.is_zero(Q[2]):
            return P
        Z1Z1 = F.mul(P[2], P[2])
        Z2Z2 = F.mul(Q[2], Q[2])
        U1 = F.mul(P[0], Z2Z2)
        U2 = F.mul(Q[0], Z1Z1)
        S1 = F.mul(P[1], F.mul(Q[2], Z2Z2))
        S2 = F.mul(Q[1], F.mul(P[2], Z1Z1))
        if U1 == U2:
            if S1 == S2:
                return self.j_double(P)
            return (F.one, F.one, F.zero)
        H = F.sub(U2, U1)
        R = F.sub(S2, S1)
        HH = F.mul(H, H)
        HHH = F.mul(H, HH)
        V = F.mul(U1, HH)
        X3 = F.sub(F.sub(F.mul(R, R), HHH), F.add(V, V))
        Y3 = F.sub(F.mul(R, F.sub(V, X3)), F.mul(S1, HHH))
        Z3 = F.mul(F.mul(P[2], Q[2]), H)
        return (X3, Y3, Z3)

    def mul(self, P, k):
        k %= self.scalar_field.p
        acc = self.j_from_affine(None)
        base = self.j_from_affine(P)
        while k:
            if k & 1:
                acc = self.j_add(acc, base)
            base = self.j_double(base)
            k >>= 1
        return self.j_to_affine(acc)

    # ---- MSM
    def msm_naive(self, points, scalars):
        acc = self.j_from_affine(None)
        for P, k in zip(points, scalars):
            k %= self.scalar_field.p
            if k == 0 or P is None:
                continue
            b = self.j_from_affine(P)
            r = self.j_from_affine(None)
            while k:
                if k & 1:
                    r = self.j_add(r, b)
                b = self.j_double(b)
                k >>= 1
            acc = self.j_add(acc, r)
        return self.j_to_affine(acc)

    def msm_bdlo12(self, points, scalars, c=None):
        """Bucket method as published in BDLO12 / libff `multi_exp_inner<multi_exp_method_BDLO12>`
        (unsigned c-bit windows, per-window buckets, running-sum reduction, c from log2 n)."""
        n = len(points)
        assert n == len(scalars)
        if n == 0:
            return None
        if c is None:
            log2n = max(1, n.bit_length() - 1)
            c = max(1, log2n - (log2n // 3 - 2)) if log2n >= 6 else max(1, log2n)
        bits = self.scalar_field.bits
        nwin = (bits + c - 1) // c
        ks = [k % self.scalar_field.p for k in scalars]
        inf = self.j_from_affine(None)
        result = inf
        for w in range(nwin - 1, -1, -1):
            for _ in range(c):
                result = self.j_double(result)
            buckets = [inf] * (1 << c)
            for P, k in zip(points, ks):
                d = (k >> (w * c)) & ((1 << c) - 1)
                if d and P is not None:
                    buckets[d] = self.j_add(buckets[d], self.j_from_affine(P))
            running = inf
            for d in range((1 << c) - 1, 0, -1):
                running = self.j_add(running, buckets[d])
                result = self.j_add(result, running)
        return self.j_to_affine(result)

    def msm_with_mixed_addition(self, points, scalars):
        """multiexp_with_mixed_addition semantics (Appendix A.4; kc variant in
        knowledge_commitment_multiexp.hpp:86-101): skip 0, add directly on 1, MSM the rest."""
        acc = None
        rp, rs = [], []
        for P, k in zip(points, scalars):
            k %= self.scalar_field.p
            if k == 0:
                continue
            if k == 1:
                acc = self.add(acc, P)
            else:
                rp.append(P)
                rs.append(k)
        return self.add(acc, self.msm_bdlo12(rp, rs) if rp else None)

    def random_points(self, n, seed):
        """Distinct synthetic points k_i*G via a running add chain (fast, deterministic)."""
        import random
        rnd = random.Random(seed)
        k0 = rnd.randrange(1, self.scalar_field.p)
        step = self.mul(self.gen, rnd.randrange(1, self.scalar_field.p))
        cur = self.j_from_affine(self.mul(self.gen, k0))
        stepj = self.j_from_affine(step)
        js = []
        for _ in range(n):
            js.append(cur)
            cur = self.j_add(cur, stepj)
        return self.batch_to_affine(js)

    def batch_to_affine(self, js):
        F = self.F
        prods, acc = [], F.one
        for J in js:
            prods.append(acc)
            if not F.is_zero(J[2]):
                acc = F.mul(acc, J[2])
        inv = F.inv(acc)
        out = [None] * len(js)
        for i in range(len(js) - 1, -1, -1):
            J = js[i]
            if F.is_zero(J[2]):
                continue
            zi = F.mul(inv, prods[i])
            inv = F.mul(inv, J[2])
            zi2 = F.mul(zi, zi)
            out[i] = (F.mul(J[0], zi2), F.mul(J[1], F.mul(zi2, zi)))
        return out


_q381 = BLS12_381_FQ.p
_q254 = BN254_FQ.p

BLS12_381_G1 = Curve(
    "bls12_381_g1", 0, FqOps(_q381), 4,
    (0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb,
     0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1),
    BLS12_381_FR, BLS12_381_FQ, 12)
BN254_G1 = Curve("bn254_g1", 1, FqOps(_q254), 3, (1, 2), BN254_FR, BN254_FQ, 8)
PALLAS = Curve("pallas", 2, FqOps(PALLAS_FP.p), 5, (PALLAS_FP.p - 1, 2), PALLAS_FQ, PALLAS_FP, 8)
BLS12_381_G2 = Curve(
    "bls12_381_g2", 3, Fq2Ops(_q381), (4, 4),
    ((0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8,
      0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e),
     (0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801,
      0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be)),
    BLS12_381_FR, BLS12_381_FQ, 24)
# BN254 G2: y^2 = x^3 + 3/(9+u)
_bn_f2 = Fq2Ops(_q254)
_bn_b2 = _bn_f2.mul((3, 0), _bn_f2.inv((9, 1)))
BN254_G2 = Curve(
    "bn254_g2", 4, _bn_f2, _bn_b2,
    ((10857046999023057135944570762232829481370756359578518086990519993285655852781,
      11559732032986387107991004021392285783925812861821192530917403151452391805634),
     (8495653923123431417604973247489272438418190587263600148770280649306958101930,
      4082367875863433681332203403145435568316851327593401208105741076214120093531)),
    BN254_FR, BN254_FQ, 16)

CURVES = {c.name: c for c in (BLS12_381_G1, BN254_G1, PALLAS, BLS12_381_G2, BN254_G2)}
CURVES_BY_ID = {c.cid: c for c in CURVES.values()}
