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
    // Operand scanning, one b-limb per row, with the running total split into two accumulators of
    // 64-bit words: `lo` holds limb positions (k, k+1), (k+2, k+3), .. and `hi` positions (k+1, k+2), ..
    // so every 32x32 partial product lands on one aligned word and each half-row is one carry chain of
    // IMAD.WIDE.U32.X.  After a row the lowest limb is zero, the roles of the accumulators swap (that is
    // the >> 32), and the surviving half of the dropped word is folded in with one 32-bit add whose
    // carry starts the next chain.
   private:
    static constexpr int H = N / 2;
    ZKB_HD static uint32_t lo32(uint64_t x) { return (uint32_t)x; }
    ZKB_HD static uint32_t hi32(uint64_t x) { return (uint32_t)(x >> 32); }
    ZKB_HD static constexpr bool is_pow2(uint32_t v) { return v > 1 && (v & (v - 1)) == 0; }
    ZKB_HD static constexpr int log2u(uint32_t v) { return v <= 1 ? 0 : 1 + log2u(v >> 1); }

    // acc[j] += p[OFF + 2j] * mi over one carry chain.  The modulus limbs are compile-time constants:
    // limbs 0, 1, 2^32-1 and 2^s (Pallas/Vesta p = 2^254 + t, t < 2^126; BLS12-381 Fr p = ..ffffffff00000001)
    // need no multiplier - their products are formed on the ALU pipe, which idles while an IMAD.WIDE
    // occupies the fma pipe for four issue cycles.  That leaves 3 (Pallas) or 6 (BLS12-381 Fr) wide
    // multiply-adds per row for the reduction instead of 8.
    template <int OFF, int J>
    ZKB_HD static void mad_row_mod_step(uint64_t *acc, uint32_t mi) {
        if constexpr (J < H) {
            constexpr uint32_t v = P::mod(OFF + 2 * J);
            if constexpr (v == 0u || v == 1u || v == 0xffffffffu || is_pow2(v)) {
                uint64_t t;
                if constexpr (v == 0u) t = 0;
                else if constexpr (v == 1u) t = ptx::pack64(mi, 0);
                else if constexpr (v == 0xffffffffu) t = ptx::pack64(0u - mi, mi - (mi != 0u ? 1u : 0u));
                else t = ptx::pack64(mi << log2u(v), mi >> (32 - log2u(v)));
                acc[J] = J == 0 ? ptx::add_cc64(acc[J], t) : ptx::addc_cc64(acc[J], t);
            } else {
                acc[J] = J == 0 ? ptx::mad_wide_cc(mi, v, acc[J]) : ptx::madc_wide_cc(mi, v, acc[J]);
            }
            mad_row_mod_step<OFF, J + 1>(acc, mi);
        }
    }
    // one row: T += a * bi; T += mi * p with mi chosen so that the lowest limb cancels
    ZKB_HD static void mad_redc(uint64_t *lo, uint64_t *hi, const uint32_t *a, uint32_t bi, bool first) {
        if (first) {
#pragma unroll
            for (int j = 0; j < H; j++) hi[j] = ptx::mul_wide(a[2 * j + 1], bi);
#pragma unroll
            for (int j = 0; j < H; j++) lo[j] = ptx::mul_wide(a[2 * j], bi);
        } else {
            lo[0] = ptx::pack64(ptx::add_cc(lo32(lo[0]), hi32(hi[0])), hi32(lo[0]));
#pragma unroll
            for (int j = 0; j < H - 1; j++) hi[j] = ptx::madc_wide_cc(a[2 * j + 1], bi, hi[j + 1]);
            hi[H - 1] = ptx::madc_wide(a[N - 1], bi, 0);
            lo[0] = ptx::mad_wide_cc(a[0], bi, lo[0]);
#pragma unroll
            for (int j = 1; j < H; j++) lo[j] = ptx::madc_wide_cc(a[2 * j], bi, lo[j]);
            hi[H - 1] = ptx::pack64(lo32(hi[H - 1]), ptx::addc(hi32(hi[H - 1]), 0));
        }
        uint32_t mi = lo32(lo[0]) * P::NINV;
        mad_row_mod_step<1, 0>(hi, mi);
        mad_row_mod_step<0, 0>(lo, mi);
        hi[H - 1] = ptx::pack64(lo32(hi[H - 1]), ptx::addc(hi32(hi[H - 1]), 0));
    }

   public:
    // Translation units whose kernels chain dozens of multiplications per loop iteration (MSM bucket
    // accumulation: ~10 per point) define ZKB_MUL_OUTLINE: one shared out-of-line copy of the body keeps
    // the loop inside the instruction cache instead of streaming ~70 KB of unrolled code per iteration.
#if defined(ZKB_MUL_OUTLINE) && defined(__CUDA_ARCH__)
    __device__ __noinline__ static Fp mul_outlined(const Fp a, const Fp b) { return mul_inline(a, b); }
    __device__ __forceinline__ friend Fp operator*(const Fp &a, const Fp &b) { return mul_outlined(a, b); }
#else
    ZKB_HD friend Fp operator*(const Fp &a, const Fp &b) { return mul_inline(a, b); }
#endif
    ZKB_HD static Fp mul_inline(const Fp &a, const Fp &b) {
        uint64_t even[H], odd[H];
#pragma unroll
        for (int i = 0; i < N; i += 2) {
            mad_redc(even, odd, a.l, b.l[i], i == 0);
            mad_redc(odd, even, a.l, b.l[i + 1], false);
        }
        // N is even, so the last row left T = odd + (even << 32) with the lowest limb of odd zero:
        // result = T >> 32 = even + (odd >> 32)
        Fp r;
        r.l[0] = ptx::add_cc(lo32(even[0]), hi32(odd[0]));
#pragma unroll
        for (int i = 1; i < N - 1; i++)
            r.l[i] = ptx::addc_cc((i & 1) ? hi32(even[i / 2]) : lo32(even[i / 2]),
                                  (i & 1) ? lo32(odd[(i + 1) / 2]) : hi32(odd[(i + 1) / 2]));
        r.l[N - 1] = ptx::addc(hi32(even[H - 1]), 0);
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

// Quadratic extension B[u]/(u^2 + 1) - the coordinate field of the G2 groups of BLS12-381 and BN254
// (both use the non-residue -1).  Generic over the base type so that the device Fp<P> and the host
// HostFp<P> share it; interface = what the group formulas of zkb_curve.cuh use.  Limb layout at the
// ABI: c0 || c1.
template <class B>
struct Fp2 {
    static constexpr int N = 2 * B::N;
    B c0, c1;

    ZKB_HD static Fp2 zero() { Fp2 r; r.c0 = B::zero(); r.c1 = B::zero(); return r; }
    ZKB_HD static Fp2 one() { Fp2 r; r.c0 = B::one(); r.c1 = B::zero(); return r; }
    ZKB_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    ZKB_HD bool operator==(const Fp2 &o) const { return c0 == o.c0 && c1 == o.c1; }
    ZKB_HD bool operator!=(const Fp2 &o) const { return !(*this == o); }
    ZKB_HD friend Fp2 operator+(const Fp2 &a, const Fp2 &b) { Fp2 r; r.c0 = a.c0 + b.c0; r.c1 = a.c1 + b.c1; return r; }
    ZKB_HD friend Fp2 operator-(const Fp2 &a, const Fp2 &b) { Fp2 r; r.c0 = a.c0 - b.c0; r.c1 = a.c1 - b.c1; return r; }
    ZKB_HD Fp2 neg() const { Fp2 r; r.c0 = c0.neg(); r.c1 = c1.neg(); return r; }
    ZKB_HD Fp2 dbl() const { Fp2 r; r.c0 = c0.dbl(); r.c1 = c1.dbl(); return r; }
    // Karatsuba: 3 base multiplications
    ZKB_HD friend Fp2 operator*(const Fp2 &a, const Fp2 &b) {
        B t0 = a.c0 * b.c0, t1 = a.c1 * b.c1;
        B t2 = (a.c0 + a.c1) * (b.c0 + b.c1);
        Fp2 r;
        r.c0 = t0 - t1;
        r.c1 = t2 - t0 - t1;
        return r;
    }
    // (a0 + a1)(a0 - a1) + 2 a0 a1 u : 2 base multiplications
    ZKB_HD Fp2 sqr() const {
        B s = c0 + c1, d = c0 - c1, m = c0 * c1;
        Fp2 r;
        r.c0 = s * d;
        r.c1 = m.dbl();
        return r;
    }
    ZKB_HD Fp2 to_mont() const { Fp2 r; r.c0 = c0.to_mont(); r.c1 = c1.to_mont(); return r; }
    ZKB_HD Fp2 from_mont() const { Fp2 r; r.c0 = c0.from_mont(); r.c1 = c1.from_mont(); return r; }
    ZKB_HD Fp2 inverse() const {   // conj / norm (0 -> 0)
        B n = (c0.sqr() + c1.sqr()).inverse();
        Fp2 r;
        r.c0 = c0 * n;
        r.c1 = (c1 * n).neg();
        return r;
    }
    ZKB_HD Fp2 &operator+=(const Fp2 &o) { *this = *this + o; return *this; }
    ZKB_HD Fp2 &operator-=(const Fp2 &o) { *this = *this - o; return *this; }
    ZKB_HD Fp2 &operator*=(const Fp2 &o) { *this = *this * o; return *this; }
};

}  // namespace zkb
