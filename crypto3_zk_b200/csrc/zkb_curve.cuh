// Short-Weierstrass y^2 = x^3 + b (a = 0) group arithmetic in XYZZ coordinates
// (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2; infinity <=> ZZ == 0).
//
// Serves the device side of `algebra::multiexp` (call sites: zk/commitments/polynomial/kzg.hpp:146,
// zk/snark/systems/ppzksnark/r1cs_gg_ppzksnark/prover.hpp:108-139).  The reference's group type is
// Jacobian; results are compared in affine form, which is representation independent.
// Formulas: Explicit-Formulas Database "xyzz" (madd-2008-s 8M+2S, add-2008-s 12M+2S,
// dbl-2008-s-1, mdbl-2008-s-1), with the exceptional cases (equal / opposite / infinite inputs)
// handled explicitly because adversarial inputs (repeated points) are part of the parity tests.
#pragma once
#include "zkb_field.cuh"

namespace zkb {

template <class F>
struct Affine {
    F x, y;  // (0,0) encodes infinity (never on a curve with b != 0)
    ZKB_HD bool is_infinity() const { return x.is_zero() && y.is_zero(); }
    ZKB_HD static Affine infinity() { Affine r; r.x = F::zero(); r.y = F::zero(); return r; }
    ZKB_HD Affine to_mont() const { Affine r; r.x = x.to_mont(); r.y = y.to_mont(); return r; }
    ZKB_HD Affine from_mont() const { Affine r; r.x = x.from_mont(); r.y = y.from_mont(); return r; }
    ZKB_HD Affine neg() const { Affine r; r.x = x; r.y = y.neg(); return r; }
};

template <class F>
struct XYZZ {
    F X, Y, ZZ, ZZZ;

    ZKB_HD static XYZZ infinity() {
        XYZZ r;
        r.X = F::zero(); r.Y = F::zero(); r.ZZ = F::zero(); r.ZZZ = F::zero();
        return r;
    }
    ZKB_HD bool is_infinity() const { return ZZ.is_zero(); }

    ZKB_HD static XYZZ from_affine(const Affine<F> &p) {
        if (p.is_infinity()) return infinity();
        XYZZ r;
        r.X = p.x; r.Y = p.y; r.ZZ = F::one(); r.ZZZ = F::one();
        return r;
    }

    // 2 * (affine p)   (mdbl-2008-s-1)
    ZKB_HD_NOINLINE static XYZZ dbl_affine(const Affine<F> p) {
        if (p.is_infinity() || p.y.is_zero()) return infinity();
        XYZZ r;
        F U = p.y.dbl();
        F V = U.sqr();
        F W = U * V;
        F S = p.x * V;
        F xx = p.x.sqr();
        F M = xx.dbl() + xx;
        r.X = M.sqr() - S.dbl();
        r.Y = M * (S - r.X) - W * p.y;
        r.ZZ = V;
        r.ZZZ = W;
        return r;
    }

    // dbl-2008-s-1
    ZKB_HD_NOINLINE XYZZ dbl() const {
        if (is_infinity() || Y.is_zero()) return infinity();
        XYZZ r;
        F U = Y.dbl();
        F V = U.sqr();
        F W = U * V;
        F S = X * V;
        F xx = X.sqr();
        F M = xx.dbl() + xx;
        r.X = M.sqr() - S.dbl();
        r.Y = M * (S - r.X) - W * Y;
        r.ZZ = V * ZZ;
        r.ZZZ = W * ZZZ;
        return r;
    }

    // this += affine p   (madd-2008-s)
    ZKB_HD void add_mixed(const Affine<F> &p) {
        if (p.is_infinity()) return;
        if (is_infinity()) {
            X = p.x; Y = p.y; ZZ = F::one(); ZZZ = F::one();
            return;
        }
        F U2 = p.x * ZZ;
        F S2 = p.y * ZZZ;
        F Pd = U2 - X;
        F R = S2 - Y;
        if (Pd.is_zero()) {
            if (R.is_zero()) *this = dbl_affine(p);
            else *this = infinity();
            return;
        }
        F PP = Pd.sqr();
        F PPP = Pd * PP;
        F Q = X * PP;
        F X3 = R.sqr() - PPP - Q.dbl();
        Y = R * (Q - X3) - Y * PPP;
        X = X3;
        ZZ = ZZ * PP;
        ZZZ = ZZZ * PPP;
    }

    // this += q   (add-2008-s)
    ZKB_HD_NOINLINE void add(const XYZZ &q) {
        if (q.is_infinity()) return;
        if (is_infinity()) { *this = q; return; }
        F U1 = X * q.ZZ;
        F U2 = q.X * ZZ;
        F S1 = Y * q.ZZZ;
        F S2 = q.Y * ZZZ;
        F Pd = U2 - U1;
        F R = S2 - S1;
        if (Pd.is_zero()) {
            if (R.is_zero()) *this = dbl();
            else *this = infinity();
            return;
        }
        F PP = Pd.sqr();
        F PPP = Pd * PP;
        F Q = U1 * PP;
        F X3 = R.sqr() - PPP - Q.dbl();
        Y = R * (Q - X3) - S1 * PPP;
        X = X3;
        ZZ = ZZ * q.ZZ * PP;
        ZZZ = ZZZ * q.ZZZ * PPP;
    }

    ZKB_HD XYZZ neg() const { XYZZ r = *this; r.Y = Y.neg(); return r; }

    // one inversion; Montgomery in, Montgomery out
    ZKB_HD Affine<F> to_affine() const {
        if (is_infinity()) return Affine<F>::infinity();
        F ti = (ZZ * ZZZ).inverse();
        Affine<F> r;
        r.x = X * (ZZZ * ti);
        r.y = Y * (ZZ * ti);
        return r;
    }

    // k * this for a small unsigned k (double-and-add, MSB first)
    ZKB_HD XYZZ mul_small(uint32_t k) const {
        XYZZ r = infinity();
        for (int b = 31; b >= 0; b--) {
            r = r.dbl();
            if ((k >> b) & 1) r.add(*this);
        }
        return r;
    }
};

}  // namespace zkb
