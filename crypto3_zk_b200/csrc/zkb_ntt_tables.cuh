// Definitions of the NTT twiddle / scaling tables (Montgomery form), shared by the device table
// generation kernels (zkb_ntt.cu) and the CPU replay harness.
//   out[idx] = scale * base^e(idx),  e(idx) = idx                      (1-D: coset powers, master tw)
//                                    e(idx) = (idx >> log_m) * (idx mod 2^log_m)   (2-D: inter-pass T_i)
#pragma once
#include "zkb_field.cuh"

namespace zkb {

template <class F>
ZKB_HD F powtab_value(const F &base, const F &scale, int two_d, int log_m, uint64_t idx) {
    uint64_t e = two_d ? (idx >> log_m) * (idx & ((1ull << log_m) - 1)) : idx;
    return scale * base.pow_u64(e);
}

// w_{2^log_n} (Montgomery) = root_of_unity^(2^(S - log_n)); inverse if inv
template <class F, class P>
ZKB_HD F ntt_omega(int log_n, bool inv) {
    F w = F::root_of_unity();
    for (int i = 0; i < P::TWO_ADICITY - log_n; i++) w = w.sqr();
    return inv ? w.inverse() : w;
}

// (2^log_n)^-1 in Montgomery form
template <class F>
ZKB_HD F ntt_n_inv(int log_n) {
    F r = F::one();
    F h = F::two_inv();
    for (int i = 0; i < log_n; i++) r = r * h;
    return r;
}

}  // namespace zkb
