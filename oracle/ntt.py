"""Radix-2 evaluation domain + polynomial_dfs restatement (oracle; test infrastructure only).

Follows crypto3-math's `basic_radix2_domain` / `polynomial_dfs` semantics (un-vendored, libfqfft
lineage) as used at the reference call sites:
  fft / inverse_fft / multiply_by_coset  r1cs_to_qap.hpp:250-315
  resize (= iFFT, zero-pad, FFT)         basic_fri.hpp:369-371,451-455
  get_domain_element                     basic_fri.hpp:783, fold_polynomial.hpp:83
  calculate_domain_set                   basic_fri.hpp:162 ; pinned by test/commitment/fri.cpp:122-123
The algorithm is the textbook in-place bit-reversal + DIT butterflies; outputs are the DFT
a_hat[i] = sum_j a[j] w^(ij) in natural order, so any correct transform must agree with it.
"""
from .fields import Field


def bitrev(i, bits):
    r = 0
    for _ in range(bits):
        r = (r << 1) | (i & 1)
        i >>= 1
    return r


def _radix2_inplace(a, omega, p):
    n = len(a)
    logn = n.bit_length() - 1
    assert 1 << logn == n
    for k in range(n):
        rk = bitrev(k, logn)
        if k < rk:
            a[k], a[rk] = a[rk], a[k]
    m = 1
    for s in range(1, logn + 1):
        w_m = pow(omega, n // (2 * m), p)
        for k in range(0, n, 2 * m):
            w = 1
            for j in range(m):
                t = w * a[k + j + m] % p
                u = a[k + j]
                a[k + j] = (u + t) % p
                a[k + j + m] = (u - t) % p
                w = w * w_m % p
        m *= 2


class EvaluationDomain:
    """math::evaluation_domain<F> for m = 2^k (basic_radix2_domain)."""

    def __init__(self, field: Field, m: int):
        if m < 1 or m & (m - 1):
            raise ValueError("only power-of-two domains (basic_radix2_domain)")
        self.field = field
        self.m = m
        self.log_m = m.bit_length() - 1
        self.omega = field.omega(self.log_m)
        self.omega_inv = field.inv(self.omega)

    def size(self):
        return self.m

    def fft(self, a):
        """In place; a shorter than m is zero-padded (upstream behaviour)."""
        if len(a) > self.m:
            raise ValueError("vector larger than domain")
        a.extend([0] * (self.m - len(a)))
        _radix2_inplace(a, self.omega, self.field.p)

    def inverse_fft(self, a):
        if len(a) > self.m:
            raise ValueError("vector larger than domain")
        a.extend([0] * (self.m - len(a)))
        p = self.field.p
        _radix2_inplace(a, self.omega_inv, p)
        minv = self.field.inv(self.m % p)
        for i in range(self.m):
            a[i] = a[i] * minv % p

    def get_domain_element(self, i):
        return pow(self.omega, i, self.field.p)

    def compute_vanishing_polynomial(self, t):
        return (pow(t, self.m, self.field.p) - 1) % self.field.p

    def add_poly_z(self, c, H):
        p = self.field.p
        H[self.m] = (H[self.m] + c) % p
        H[0] = (H[0] - c) % p

    def divide_by_z_on_coset(self, P):
        p = self.field.p
        zinv = self.field.inv(self.compute_vanishing_polynomial(self.field.g))
        for i in range(self.m):
            P[i] = P[i] * zinv % p

    def evaluate_all_lagrange_polynomials(self, t):
        p, m = self.field.p, self.m
        if m == 1:
            return [1]
        if pow(t, m, p) == 1:
            u = [0] * m
            w = 1
            for i in range(m):
                if w == t % p:
                    u[i] = 1
                    return u
                w = w * self.omega % p
        z = (pow(t, m, p) - 1) % p
        l = z * self.field.inv(m) % p
        r = 1
        u = []
        for i in range(m):
            u.append(l * self.field.inv((t - r) % p) % p)
            l = l * self.omega % p
            r = r * self.omega % p
        return u


def multiply_by_coset(a, g, p):
    """math::multiply_by_coset: a[i] *= g^i (r1cs_to_qap.hpp:266)."""
    u = 1
    for i in range(len(a)):
        a[i] = a[i] * u % p
        u = u * g % p


def make_evaluation_domain(field, m):
    """Only the power-of-two branch of upstream make_evaluation_domain is restated; other m round
    up is NOT upstream behaviour (it picks extended/step radix-2) so we refuse instead."""
    return EvaluationDomain(field, m)


def calculate_domain_set(field, max_log, set_size):
    return [EvaluationDomain(field, 1 << (max_log - i)) for i in range(set_size)]


def dft_naive(a, omega, p):
    n = len(a)
    return [sum(a[j] * pow(omega, i * j, p) for j in range(n)) % p for i in range(n)]


# --------------------------------------------------------------------------- polynomial_dfs pieces
def dfs_resize(vals, field, new_size):
    """polynomial_dfs::resize(sz, nullptr, D) (Appendix A.3): size()==1 replicates the constant,
    else inverse_fft on the own-size domain, zero-pad (or truncate), fft on the new domain."""
    if len(vals) == new_size:
        return list(vals)
    if len(vals) == 1:
        return [vals[0]] * new_size
    c = list(vals)
    EvaluationDomain(field, len(c)).inverse_fft(c)
    if new_size < len(c):
        c = c[:new_size]
    EvaluationDomain(field, new_size).fft(c)
    return c


def dfs_coefficients(vals, field):
    c = list(vals)
    EvaluationDomain(field, len(c)).inverse_fft(c)
    return c


def dfs_from_coefficients(coeffs, field):
    n = 1
    while n < len(coeffs):
        n *= 2
    c = list(coeffs)
    EvaluationDomain(field, n).fft(c)
    return c


def dfs_evaluate(vals, field, x):
    c = dfs_coefficients(vals, field)
    r = 0
    for v in reversed(c):
        r = (r * x + v) % field.p
    return r


def coset_fft(a, field, shift):
    """multiply_by_coset(a, g) then fft(a) (r1cs_to_qap.hpp:266-270)."""
    a = list(a)
    multiply_by_coset(a, shift, field.p)
    EvaluationDomain(field, len(a)).fft(a)
    return a


def coset_inverse_fft(a, field, shift):
    """inverse_fft(a) then multiply_by_coset(a, g^-1) (r1cs_to_qap.hpp:308-315)."""
    a = list(a)
    EvaluationDomain(field, len(a)).inverse_fft(a)
    multiply_by_coset(a, field.inv(shift), field.p)
    return a
