"""Groth16 (r1cs_gg_ppzksnark) restatement on plain Python integers (oracle; test infrastructure only).

In-repo, exact:
  R1CS example with field input     test/systems/ppzksnark/r1cs_examples.hpp:77-146
  R1CS example with binary input    test/systems/ppzksnark/r1cs_examples.hpp:156-212
  swap_AB_if_beneficial             zk/snark/arithmetization/constraint_satisfaction_problems/r1cs.hpp:193-218
  instance_map_with_evaluation      zk/snark/reductions/r1cs_to_qap.hpp:137-184
  witness_map                       zk/snark/reductions/r1cs_to_qap.hpp:219-325
  generator (deterministic form)    zk/snark/systems/ppzksnark/r1cs_gg_ppzksnark/generator.hpp:86-237
  prover                            zk/snark/systems/ppzksnark/r1cs_gg_ppzksnark/prover.hpp:73-158
Group results are compared in affine form, so they do not depend on the MSM algorithm; the restatement is pinned
by the Groth16 relations in the exponent (tests/test_oracle_groth16.py): with the toxic waste known,
A = (alpha + sum x_i A_i(t) + r delta) G1 and so on.
"""
import random

from . import ntt


class R1cs:
    """constraints: list of (a, b, c), each a list of (variable index, coefficient); index 0 is the constant 1."""

    def __init__(self, num_inputs, num_aux, constraints):
        self.num_inputs, self.num_aux, self.constraints = num_inputs, num_aux, constraints

    @property
    def num_variables(self):
        return self.num_inputs + self.num_aux

    @property
    def num_constraints(self):
        return len(self.constraints)

    def swap_ab_if_beneficial(self):
        ta, tb = set(), set()
        for a, b, _ in self.constraints:
            ta.update(i for i, _ in a)
            tb.update(i for i, _ in b)
        if len(tb) > len(ta):
            self.constraints = [(b, a, c) for a, b, c in self.constraints]

    def is_satisfied(self, full, p):
        x = [1] + list(full)
        ev = lambda lc: sum(co * x[i] for i, co in lc) % p
        return all(ev(a) * ev(b) % p == ev(c) for a, b, c in self.constraints)


def example_with_field_input(field, num_constraints, num_inputs, seed=0):
    p = field.p
    rnd = random.Random(seed)
    a, b = rnd.randrange(p), rnd.randrange(p)
    full = [a, b]
    cons = []
    for i in range(num_constraints - 1):
        if i % 2:
            cons.append(([(i + 1, 1)], [(i + 2, 1)], [(i + 3, 1)]))
            tmp = a * b % p
        else:
            cons.append(([(i + 1, 1), (i + 2, 1)], [(0, 1)], [(i + 3, 1)]))
            tmp = (a + b) % p
        full.append(tmp)
        a, b = b, tmp
    num_vars = 2 + num_constraints     # primary + auxiliary = num_inputs + (2 + num_constraints - num_inputs)
    fin = sum(full[i - 1] for i in range(1, num_vars)) % p
    cons.append(([(i, 1) for i in range(1, num_vars)], [(i, 1) for i in range(1, num_vars)], [(num_vars, 1)]))
    full.append(fin * fin % p)
    cs = R1cs(num_inputs, 2 + num_constraints - num_inputs, cons)
    assert cs.num_variables == len(full) and cs.is_satisfied(full, p)
    return cs, full[:num_inputs], full[num_inputs:]


def example_with_binary_input(field, num_constraints, num_inputs, seed=0):
    p = field.p
    rnd = random.Random(seed)
    full = [rnd.randrange(2) for _ in range(num_inputs)]
    cons = []
    lastvar = num_inputs - 1
    for i in range(num_constraints):
        lastvar += 1
        u = rnd.randrange(num_inputs) if i == 0 else rnd.randrange(i)
        v = rnd.randrange(num_inputs) if i == 0 else rnd.randrange(i)
        c = [(u + 1, 2)] if u == v else [(u + 1, 1), (v + 1, 1)]
        c.append((lastvar + 1, p - 1))
        cons.append(([(u + 1, 2)], [(v + 1, 1)], c))
        full.append((full[u] + full[v] - 2 * full[u] * full[v]) % p)
    cs = R1cs(num_inputs, num_constraints, cons)
    assert cs.is_satisfied(full, p)
    return cs, full[:num_inputs], full[num_inputs:]


def instance_map_with_evaluation(cs, t, field):
    p = field.p
    dom = ntt.make_evaluation_domain(field, cs.num_constraints + cs.num_inputs + 1)
    At, Bt, Ct = ([0] * (cs.num_variables + 1) for _ in range(3))
    Zt = dom.compute_vanishing_polynomial(t)
    u = dom.evaluate_all_lagrange_polynomials(t)
    for i in range(cs.num_inputs + 1):
        At[i] = u[cs.num_constraints + i]
    for i, (a, b, c) in enumerate(cs.constraints):
        for idx, co in a:
            At[idx] = (At[idx] + u[i] * co) % p
        for idx, co in b:
            Bt[idx] = (Bt[idx] + u[i] * co) % p
        for idx, co in c:
            Ct[idx] = (Ct[idx] + u[i] * co) % p
    Ht = [pow(t, i, p) for i in range(dom.m + 1)]
    return dom, At, Bt, Ct, Ht, Zt


def witness_map(cs, primary, aux, field, d1=0, d2=0, d3=0):
    p, g = field.p, field.g
    dom = ntt.make_evaluation_domain(field, cs.num_constraints + cs.num_inputs + 1)
    m = dom.m
    full = list(primary) + list(aux)
    x = [1] + full
    ev = lambda lc: sum(co * x[i] for i, co in lc) % p
    aA, aB = [0] * m, [0] * m
    for i in range(cs.num_inputs + 1):
        aA[i + cs.num_constraints] = full[i - 1] if i > 0 else 1
    for i, (a, b, _) in enumerate(cs.constraints):
        aA[i] = (aA[i] + ev(a)) % p
        aB[i] = (aB[i] + ev(b)) % p
    dom.inverse_fft(aA)
    dom.inverse_fft(aB)
    coeffs = [(d2 * x_ + d1 * y_) % p for x_, y_ in zip(aA, aB)] + [0]
    coeffs[0] = (coeffs[0] - d3) % p
    dom.add_poly_z(d1 * d2 % p, coeffs)
    aA, aB = ntt.coset_fft(aA, field, g), ntt.coset_fft(aB, field, g)
    H = [x_ * y_ % p for x_, y_ in zip(aA, aB)]
    aC = [0] * m
    for i, (_, _, c) in enumerate(cs.constraints):
        aC[i] = ev(c)
    dom.inverse_fft(aC)
    aC = ntt.coset_fft(aC, field, g)
    H = [(h - c) % p for h, c in zip(H, aC)]
    dom.divide_by_z_on_coset(H)
    H = ntt.coset_inverse_fft(H, field, g)
    for i in range(m):
        coeffs[i] = (coeffs[i] + H[i]) % p
    return m, full, coeffs


class ProvingKey:
    pass


def generator(cs, curve_g1, curve_g2, field, t, alpha, beta, gamma, delta, g1=None, g2=None):
    """deterministic_basic_process (generator.hpp:223-350 = basic_process with the randomness passed in); also keeps
    the query scalars so that tests can check proofs in the exponent."""
    p = field.p
    g1 = curve_g1.gen if g1 is None else g1
    g2 = curve_g2.gen if g2 is None else g2
    cs = R1cs(cs.num_inputs, cs.num_aux, list(cs.constraints))
    cs.swap_ab_if_beneficial()
    dom, At, Bt, Ct, Ht, Zt = instance_map_with_evaluation(cs, t, field)
    dinv = field.inv(delta)
    off = cs.num_inputs + 1
    Lt = [(beta * At[off + i] + alpha * Bt[off + i] + Ct[off + i]) * dinv % p for i in range(cs.num_variables - cs.num_inputs)]
    Ht = Ht[:len(Ht) - 2]
    coeff = Zt * dinv % p
    pk = ProvingKey()
    pk.cs = cs
    pk.scalars = dict(At=At, Bt=Bt, Ct=Ct, Lt=Lt, Ht=[coeff * h % p for h in Ht], Zt=Zt,
                      t=t, alpha=alpha, beta=beta, gamma=gamma, delta=delta)
    pk.alpha_g1, pk.beta_g1, pk.delta_g1 = (curve_g1.mul(g1, s) for s in (alpha, beta, delta))
    pk.beta_g2, pk.delta_g2 = curve_g2.mul(g2, beta), curve_g2.mul(g2, delta)
    pk.A_query = [curve_g1.mul(g1, s) for s in At]
    # kc_batch_exp keeps only the non-zero entries (sparse_vector: indices + values)
    pk.B_indices = [i for i, s in enumerate(Bt) if s]
    pk.B_g2 = [curve_g2.mul(g2, Bt[i]) for i in pk.B_indices]
    pk.B_g1 = [curve_g1.mul(g1, Bt[i]) for i in pk.B_indices]
    pk.H_query = [curve_g1.mul(g1, s) for s in pk.scalars["Ht"]]
    pk.L_query = [curve_g1.mul(g1, s) for s in Lt]
    pk.domain_size = dom.m
    return pk


def prove(pk, primary, aux, r, s, curve_g1, curve_g2, field):
    """r1cs_gg_ppzksnark_prover::process (prover.hpp:73-158) with the zero-knowledge randomness (r, s) passed in."""
    p = field.p
    cs = pk.cs
    degree, full, H = witness_map(cs, primary, aux, field)
    assert H[degree - 1] == 0 and H[degree] == 0
    padded = [1] + full
    nv, ni = cs.num_variables, cs.num_inputs
    eval_A = curve_g1.msm_with_mixed_addition(pk.A_query[:nv + 1], padded[:nv + 1])
    sc_B = [padded[i] for i in pk.B_indices if i < nv + 1]
    eval_B_g2 = curve_g2.msm_with_mixed_addition(pk.B_g2[:len(sc_B)], sc_B)
    eval_B_g1 = curve_g1.msm_with_mixed_addition(pk.B_g1[:len(sc_B)], sc_B)
    eval_H = curve_g1.msm_naive(pk.H_query[:degree - 1], H[:degree - 1])
    eval_L = curve_g1.msm_with_mixed_addition(pk.L_query, padded[ni + 1:nv + 1])
    add1, add2 = curve_g1.add, curve_g2.add
    g1_A = add1(add1(pk.alpha_g1, eval_A), curve_g1.mul(pk.delta_g1, r))
    g1_B = add1(add1(pk.beta_g1, eval_B_g1), curve_g1.mul(pk.delta_g1, s))
    g2_B = add2(add2(pk.beta_g2, eval_B_g2), curve_g2.mul(pk.delta_g2, s))
    g1_C = add1(add1(add1(add1(eval_H, eval_L), curve_g1.mul(g1_A, s)), curve_g1.mul(g1_B, r)),
                curve_g1.neg(curve_g1.mul(pk.delta_g1, r * s % p)))
    return g1_A, g2_B, g1_C


def proof_in_the_exponent(pk, primary, aux, r, s, field):
    """Discrete logs of (A, B, C) from the toxic waste: the Groth16 equations the proof must satisfy."""
    p = field.p
    sc = pk.scalars
    x = [1] + list(primary) + list(aux)
    _, _, H = witness_map(pk.cs, primary, aux, field)
    a = (sc["alpha"] + sum(xi * ai for xi, ai in zip(x, sc["At"])) + r * sc["delta"]) % p
    b = (sc["beta"] + sum(xi * bi for xi, bi in zip(x, sc["Bt"])) + s * sc["delta"]) % p
    ni = pk.cs.num_inputs
    c = (sum(h * q for h, q in zip(H, sc["Ht"])) + sum(xi * li for xi, li in zip(x[ni + 1:], sc["Lt"]))
         + s * a + r * b - r * s * sc["delta"]) % p
    return a, b, c
