// Prime-field arithmetic on 32-bit limbs, Montgomery form R = 2^(32 N).
//
// Replaces (on the device) the modular arithmetic crypto3-zk gets from the un-vendored
// crypto3-multiprecision `modular_adaptor` for BLS12-381 Fr/Fq,
// BN254 Fr/Fq and Pallas Fp/Fq.  All six moduli leave at least one spare bit in the top limb,
// which the interleaved multiplier below relies on (no carry out of the odd accumulator).
//
// mont_mul: operand-scanning Montgomery multiplication with the partial products split into an
// "even" and an "odd" accumulator so that every (lo,hi) pair lands on its own aligned register
// pair and each row is ONE uninterrupted carry chain: ptxas emits N IMAD.WIDE.U32 per row
// (2 N^2 per multiplication) instead of 4 N^2 narrow IMADs.
#pragma once
#include "zkb_ptx.cuh"
#include "zkb_params.cuh"

namespace zkb {

template <class P>
struct Fp {
    static constexpr int N = P::N;
    uint32_t l[N];

    // ------------------------------------------------------------------ constants
    ZKB_HD static Fp zero() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = 0;
        return r;
    }
    ZKB_HD static Fp one() {  // Montgomery one
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::r1(i);
        return r;
    }
    ZKB_HD static Fp modulus() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::mod(i);
        return r;
    }
    ZKB_HD static Fp r2() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::r2(i);
        return r;
    }
    ZKB_HD static Fp r3() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::r3(i);
        return r;
    }
    ZKB_HD static Fp two_inv() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::two_inv_mont(i);
        return r;
    }
    ZKB_HD static Fp generator() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::gen_mont(i);
        return r;
    }
    ZKB_HD static Fp root_of_unity() {
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = P::rou_mont(i);
        return r;
    }

    // ------------------------------------------------------------------ predicates
    ZKB_HD bool is_zero() const {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) acc |= l[i];
        return acc == 0;
    }
    ZKB_HD bool operator==(const Fp &o) const {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < N; i++) acc |= l[i] ^ o.l[i];
        return acc == 0;
    }
    ZKB_HD bool operator!=(const Fp &o) const { return !(*this == o); }

    // ------------------------------------------------------------------ add / sub
    // r = a - p if a >= p else a   (a < 2p)
    ZKB_HD static Fp reduce_once(const Fp &a) {
        Fp t;
        t.l[0] = ptx::sub_cc(a.l[0], P::mod(0));
#pragma unroll
        for (int i = 1; i < N; i++) t.l[i] = ptx::subc_cc(a.l[i], P::mod(i));
        uint32_t borrow = ptx::subc(0, 0);  // 0xffffffff if a < p
        Fp r;
#pragma unroll
        for (int i = 0; i < N; i++) r.l[i] = borrow ? a.l[i] : t.l[i];
        return r;
    }

    ZKB_HD friend Fp operator+(const Fp &a, const Fp &b) {
        Fp s;
        s.l[0] = ptx::add_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) s.l[i] = ptx::addc_cc(a.l[i], b.l[i]);
        s.l[N - 1] = ptx::addc(a.l[N - 1], b.l[N - 1]);  // spare bit: no carry out
        return reduce_once(s);
    }

    ZKB_HD friend Fp operator-(const Fp &a, const Fp &b) {
        Fp d;
        d.l[0] = ptx::sub_cc(a.l[0], b.l[0]);
#pragma unroll
        for (int i = 1; i < N; i++) d.l[i] = ptx::subc_cc(a.l[i], b.l[i]);
        uint32_t borrow = ptx::subc(0, 0);  // all-ones when a < b
        Fp r;
        r.l[0] = ptx::add_cc(d.l[0], P::mod(0) & borrow);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(d.l[i], P::mod(i) & borrow);
        r.l[N - 1] = ptx::addc(d.l[N - 1], P::mod(N - 1) & borrow);
        return r;
    }

    ZKB_HD Fp neg() const {
        if (is_zero()) return *this;
        Fp r;
        r.l[0] = ptx::sub_cc(P::mod(0), l[0]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = ptx::subc_cc(P::mod(i), l[i]);
        r.l[N - 1] = ptx::subc(P::mod(N - 1), l[N - 1]);
        return r;
    }
    ZKB_HD Fp dbl() const { return *this + *this; }

    // ------------------------------------------------------------------ Montgomery multiplication
   private:
    // acc[j], acc[j+1] = a[j] * bi  for j = 0, 2, ..  (a points at the even- or odd-indexed limbs)
    ZKB_HD static void mul_row(uint32_t *acc, const uint32_t *a, uint32_t bi) {
#pragma unroll
        for (int j = 0; j < N; j += 2) {
            acc[j] = ptx::mul_lo(a[j], bi);
            acc[j + 1] = ptx::mul_hi(a[j], bi);
        }
    }
    // acc += {a[j] * bi}, one carry chain over the whole row; carry-out stays in CC
    ZKB_HD static void mad_row(uint32_t *acc, const uint32_t *a, uint32_t bi) {
        acc[0] = ptx::mad_lo_cc(a[0], bi, acc[0]);
        acc[1] = ptx::madc_hi_cc(a[0], bi, acc[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            acc[j] = ptx::madc_lo_cc(a[j], bi, acc[j]);
            acc[j + 1] = ptx::madc_hi_cc(a[j], bi, acc[j + 1]);
        }
    }
    // same with the modulus as the multiplicand (compile-time constants -> immediates)
    template <int OFF>
    ZKB_HD static void mad_row_mod(uint32_t *acc, uint32_t mi) {
        acc[0] = ptx::mad_lo_cc(P::mod(OFF), mi, acc[0]);
        acc[1] = ptx::madc_hi_cc(P::mod(OFF), mi, acc[1]);
#pragma unroll
        for (int j = 2; j < N; j += 2) {
            acc[j] = ptx::madc_lo_cc(P::mod(OFF + j), mi, acc[j]);
            acc[j + 1] = ptx::madc_hi_cc(P::mod(OFF + j), mi, acc[j + 1]);
        }
    }
    // acc = (acc >> 64) + {a[j] * bi}, continuing the carry chain already in CC
    ZKB_HD static void mad_row_shift2(uint32_t *acc, const uint32_t *a, uint32_t bi) {
#pragma unroll
        for (int j = 0; j < N - 2; j += 2) {
            acc[j] = ptx::madc_lo_cc(a[j], bi, acc[j + 2]);
            acc[j + 1] = ptx::madc_hi_cc(a[j], bi, acc[j + 3]);
        }
        acc[N - 2] = ptx::madc_lo_cc(a[N - 2], bi, 0);
        acc[N - 1] = ptx::madc_hi(a[N - 2], bi, 0);
    }
    // One outer iteration: T += a*bi; T += m*p; (the >>32 is realised by swapping lo/hi roles)
    // `lo` holds limb positions k, `hi` holds positions k+1.
    ZKB_HD static void mad_redc(uint32_t *lo, uint32_t *hi, const uint32_t *a, uint32_t bi, bool first) {
        if (first) {
            mul_row(hi, a + 1, bi);
            mul_row(lo, a, bi);
        } else {
            lo[0] = ptx::add_cc(lo[0], hi[1]);
            mad_row_shift2(hi, a + 1, bi);
            mad_row(lo, a, bi);
            hi[N - 1] = ptx::addc(hi[N - 1], 0);
        }
        uint32_t mi = lo[0] * P::NINV;
        mad_row_mod<1>(hi, mi);
        mad_row_mod<0>(lo, mi);
        hi[N - 1] = ptx::addc(hi[N - 1], 0);
    }

   public:
    ZKB_HD friend Fp operator*(const Fp &a, const Fp &b) {
        uint32_t even[N], odd[N];
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            mad_redc(even, odd, a.l, b.l[i], i == 0);
            mad_redc(odd, even, a.l, b.l[i + 1], false);
        }
        // N is even, so the last row left T = odd + (even << 32) with odd[0] == 0:
        // result = T >> 32 = even + (odd >> 32)
        Fp r;
        r.l[0] = ptx::add_cc(even[0], odd[1]);
#pragma unroll
        for (int i = 1; i < N - 1; i++) r.l[i] = ptx::addc_cc(even[i], odd[i + 1]);
        r.l[N - 1] = ptx::addc(even[N - 1], 0);
        return reduce_once(r);
    }
    ZKB_HD Fp sqr() const { return *this * *this; }

    ZKB_HD Fp &operator+=(const Fp &o) { *this = *this + o; return *this; }
    ZKB_HD Fp &operator-=(const Fp &o) { *this = *this - o; return *this; }
    ZKB_HD Fp &operator*=(const Fp &o) { *this = *this * o; return *this; }

    // ------------------------------------------------------------------ form conversion
    ZKB_HD Fp to_mont() const { return *this * r2(); }
    ZKB_HD Fp from_mont() const {
        Fp o = zero();
        o.l[0] = 1;
        return *this * o;
    }

    // ------------------------------------------------------------------ exponentiation / inverse
    // this^e for a little-endian limb exponent (not constant time; not needed here)
    ZKB_HD Fp pow_limbs(const uint32_t *e, int nlimbs) const {
        Fp r = one();
        bool started = false;
        for (int i = nlimbs - 1; i >= 0; i--) {
            for (int b = 31; b >= 0; b--) {
                if (started) r = r.sqr();
                if ((e[i] >> b) & 1) {
                    r = started ? r * *this : *this;
                    started = true;
                }
            }
        }
        return r;
    }
    ZKB_HD Fp pow_u64(uint64_t e) const {
        uint32_t ee[2] = {(uint32_t)e, (uint32_t)(e >> 32)};
        return pow_limbs(ee, 2);
    }
    // Fermat inverse (0 -> 0)
    ZKB_HD Fp inverse() const {
        uint32_t e[N];
        e[0] = ptx::sub_cc(P::mod(0), 2);
#pragma unroll
        for (int i = 1; i < N; i++) e[i] = ptx::subc_cc(P::mod(i), 0);
        return pow_limbs(e, N);
    }
};

}  // namespace zkb
