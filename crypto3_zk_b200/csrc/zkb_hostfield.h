// Fast host-side Montgomery field (64-bit limbs, unsigned __int128) with the same interface as the
// device Fp<P>, so XYZZ<HostFp<P>> reuses the group formulas of zkb_curve.cuh.  Used only for the
// O(1)-sized host tails of the MSM (window combine: ~c*W doublings, one inversion) - never as a
// substitute for a kernel.  Same R = 2^(32 N) as the device, so device XYZZ limbs are consumed as is.
#pragma once
#include <stdint.h>
#include <string.h>
#include "zkb_params.cuh"

namespace zkb {

template <class P>
struct HostFp {
    static constexpr int N = P::N;       // 32-bit limbs (ABI)
    static constexpr int M = P::N / 2;   // 64-bit limbs
    uint64_t v[M];

    static uint64_t modl(int i) { return (uint64_t)P::mod(2 * i) | ((uint64_t)P::mod(2 * i + 1) << 32); }
    static uint64_t ninv64() {  // -p^-1 mod 2^64 by Newton iteration
        uint64_t p0 = modl(0), x = 1;
        for (int i = 0; i < 6; i++) x *= 2 - p0 * x;
        return (uint64_t)0 - x;
    }
    static HostFp from_limbs32(const uint32_t *l) {
        HostFp r;
        for (int i = 0; i < M; i++) r.v[i] = (uint64_t)l[2 * i] | ((uint64_t)l[2 * i + 1] << 32);
        return r;
    }
    void to_limbs32(uint32_t *l) const {
        for (int i = 0; i < M; i++) {
            l[2 * i] = (uint32_t)v[i];
            l[2 * i + 1] = (uint32_t)(v[i] >> 32);
        }
    }
    template <class GET>
    static HostFp from_const(GET get) {
        uint32_t l[N];
        for (int i = 0; i < N; i++) l[i] = get(i);
        return from_limbs32(l);
    }
    static HostFp zero() { HostFp r; memset(r.v, 0, sizeof(r.v)); return r; }
    static HostFp one() { return from_const([](int i) { return P::r1(i); }); }
    static HostFp r2() { return from_const([](int i) { return P::r2(i); }); }

    bool is_zero() const {
        uint64_t a = 0;
        for (int i = 0; i < M; i++) a |= v[i];
        return a == 0;
    }
    bool operator==(const HostFp &o) const { return memcmp(v, o.v, sizeof(v)) == 0; }
    bool operator!=(const HostFp &o) const { return !(*this == o); }

    static bool geq_mod(const uint64_t *a) {
        for (int i = M - 1; i >= 0; i--) {
            if (a[i] > modl(i)) return true;
            if (a[i] < modl(i)) return false;
        }
        return true;
    }
    static void sub_mod(uint64_t *a) {
        unsigned __int128 br = 0;
        for (int i = 0; i < M; i++) {
            unsigned __int128 t = (unsigned __int128)a[i] - modl(i) - (uint64_t)br;
            a[i] = (uint64_t)t;
            br = (t >> 64) & 1;
        }
    }
    friend HostFp operator+(const HostFp &a, const HostFp &b) {
        HostFp r;
        unsigned __int128 c = 0;
        for (int i = 0; i < M; i++) {
            c += (unsigned __int128)a.v[i] + b.v[i];
            r.v[i] = (uint64_t)c;
            c >>= 64;
        }
        if (geq_mod(r.v)) sub_mod(r.v);   // spare bit: no carry out
        return r;
    }
    friend HostFp operator-(const HostFp &a, const HostFp &b) {
        HostFp r;
        unsigned __int128 br = 0;
        for (int i = 0; i < M; i++) {
            unsigned __int128 t = (unsigned __int128)a.v[i] - b.v[i] - (uint64_t)br;
            r.v[i] = (uint64_t)t;
            br = (t >> 64) & 1;
        }
        if (br) {
            unsigned __int128 c = 0;
            for (int i = 0; i < M; i++) {
                c += (unsigned __int128)r.v[i] + modl(i);
                r.v[i] = (uint64_t)c;
                c >>= 64;
            }
        }
        return r;
    }
    HostFp neg() const { return is_zero() ? *this : zero() - *this; }
    HostFp dbl() const { return *this + *this; }

    friend HostFp operator*(const HostFp &a, const HostFp &b) {  // CIOS
        static const uint64_t ninv = ninv64();
        uint64_t t[M + 2];
        memset(t, 0, sizeof(t));
        for (int i = 0; i < M; i++) {
            unsigned __int128 c = 0;
            for (int j = 0; j < M; j++) {
                c += (unsigned __int128)a.v[j] * b.v[i] + t[j];
                t[j] = (uint64_t)c;
                c >>= 64;
            }
            c += t[M];
            t[M] = (uint64_t)c;
            t[M + 1] = (uint64_t)(c >> 64);
            uint64_t m = t[0] * ninv;
            c = (unsigned __int128)m * modl(0) + t[0];
            c >>= 64;
            for (int j = 1; j < M; j++) {
                c += (unsigned __int128)m * modl(j) + t[j];
                t[j - 1] = (uint64_t)c;
                c >>= 64;
            }
            c += t[M];
            t[M - 1] = (uint64_t)c;
            t[M] = t[M + 1] + (uint64_t)(c >> 64);
        }
        HostFp r;
        memcpy(r.v, t, sizeof(r.v));
        if (t[M] || geq_mod(r.v)) sub_mod(r.v);
        return r;
    }
    HostFp sqr() const { return *this * *this; }
    HostFp to_mont() const { return *this * r2(); }
    HostFp from_mont() const {
        HostFp o = zero();
        o.v[0] = 1;
        return *this * o;
    }
    HostFp inverse() const {  // Fermat
        uint64_t e[M];
        for (int i = 0; i < M; i++) e[i] = modl(i);
        e[0] -= 2;  // p is odd and > 2: no borrow
        HostFp r = one();
        for (int i = M - 1; i >= 0; i--)
            for (int b = 63; b >= 0; b--) {
                r = r.sqr();
                if ((e[i] >> b) & 1) r = r * *this;
            }
        return r;
    }
};

}  // namespace zkb
