"""Host half of the product's Groth16 generator (crypto3_zk_b200/groth16.py: generator; reference generator.hpp:83-235,
r1cs_to_qap.hpp:147-204) on CPU: the QAP evaluation at t, the scalar vectors and the slicing of the batch-exponentiation
output, with the device's zkb_batch_exp stood in for by the oracle's scalar multiplication.  The device leg is
tests/test_gpu_flows.py::test_groth16_generator_vs_oracle."""
import numpy as np
import pytest

from crypto3_zk_b200 import groth16 as dg
from crypto3_zk_b200.api import _ints
from oracle import curves, fields, groth16 as og


class _OracleBatchExp:
    """ctx.batch_exp's contract: [n, 8] scalar limbs -> affine [n, 2, coord_limbs], zero scalar -> all-zero point"""

    def batch_exp(self, cname, base, arr):
        C = curves.CURVES[cname]
        cl = C.coord_limbs32
        sc = _ints(arr)
        out = np.zeros((len(sc), 2, cl), dtype=np.uint32)
        for i, k in enumerate(sc):
            P = C.mul(base, k)
            if P is None:
                continue
            for c in range(2):
                comp = P[c] if isinstance(P[c], tuple) else (P[c],)
                h = cl // len(comp)
                for j, v in enumerate(comp):
                    for l in range(h):
                        out[i, c, j * h + l] = (v >> (32 * l)) & 0xFFFFFFFF
        return out


CASES = [(fields.BLS12_381_FR, curves.BLS12_381_G1, curves.BLS12_381_G2, "field", 13, 2),
         (fields.BN254_FR, curves.BN254_G1, curves.BN254_G2, "binary", 27, 4)]


@pytest.mark.parametrize("F,G1,G2,kind,nc,ni", CASES)
def test_generator_host_half_vs_oracle(F, G1, G2, kind, nc, ni):
    make = og.example_with_field_input if kind == "field" else og.example_with_binary_input
    cs, primary, aux = make(F, nc, ni, seed=5)
    t, alpha, beta, gamma, delta = fields.random_elements(F, 5, 21)
    pk = og.generator(cs, G1, G2, F, t, alpha, beta, gamma, delta)
    pcs = dg.R1csConstraintSystem(cs.num_inputs, cs.num_aux, list(cs.constraints))
    At, Bt, Ct, Ht, Zt, m = dg.qap_instance_evaluation(dg.swap_ab_if_beneficial(pcs), dg.FIELD_BY_NAME[F.name], t)
    sc = pk.scalars
    assert (At, Bt, Ct, Zt, m) == (sc["At"], sc["Bt"], sc["Ct"], sc["Zt"], pk.domain_size)
    assert Ht == [pow(t, i, F.p) for i in range(m + 1)]
    key, vk = dg.generator(_OracleBatchExp(), G1.name, G2.name, pcs, t, alpha, beta, gamma, delta)
    for name in ("alpha_g1", "beta_g1", "beta_g2", "delta_g1", "delta_g2", "A_query", "B_indices", "B_g2", "B_g1",
                 "H_query", "L_query"):
        assert key[name] == getattr(pk, name), name
    assert [tuple(map(list, c)) for c in key["constraints"]] == [tuple(map(list, c)) for c in pk.cs.constraints]
    assert key["B_domain_size"] == cs.num_variables + 1
    ginv = F.inv(gamma)
    abc = [(beta * At[i] + alpha * Bt[i] + Ct[i]) * ginv % F.p for i in range(ni + 1)]
    assert vk["gamma_ABC_g1"] == (G1.mul(G1.gen, abc[0]), [G1.mul(G1.gen, v) for v in abc[1:]])
    assert vk["gamma_g2"] == G2.mul(G2.gen, gamma) and vk["delta_g2"] == pk.delta_g2 and vk["gamma_g1"] == G1.mul(G1.gen, gamma)
    # the oracle prover accepts the product's key: same proof as with the oracle's own key
    pk2 = og.ProvingKey()
    pk2.cs = og.R1cs(key["num_inputs"], key["num_aux"], key["constraints"])
    for name in ("alpha_g1", "beta_g1", "beta_g2", "delta_g1", "delta_g2", "A_query", "B_indices", "B_g2", "B_g1",
                 "H_query", "L_query"):
        setattr(pk2, name, key[name])
    r, s = fields.random_elements(F, 2, 22)
    assert og.prove(pk2, primary, aux, r, s, G1, G2, F) == og.prove(pk, primary, aux, r, s, G1, G2, F)


def test_lagrange_at_domain_element_and_batch_inverse():
    F = dg.FIELD_BY_NAME["bn254_fr"]
    from crypto3_zk_b200.fields import omega
    w = omega(F, 4)
    u = dg._lagrange_at(F, 4, pow(w, 3, F.p))
    assert u == [1 if i == 3 else 0 for i in range(16)]
    u = dg._lagrange_at(F, 4, 12345)
    assert sum(u) % F.p == 1                       # the Lagrange basis sums to the constant 1
    # interpolation property: sum_i L_i(t) w^(i k) = t^k for k < m
    assert sum(ui * pow(w, 5 * i, F.p) for i, ui in enumerate(u)) % F.p == pow(12345, 5, F.p)
    vals = [3, 7, F.p - 1, 123456789]
    assert [v * i % F.p for v, i in zip(vals, dg._batch_inverse(vals, F.p))] == [1] * 4


def test_generator_rejects_non_radix2_domain():
    from crypto3_zk_b200 import capi
    cs = dg.R1csConstraintSystem(1, 2, [([(0, 1)], [(1, 1)], [(2, 1)])] * 3)      # 3 + 1 + 1 = 5
    with pytest.raises(capi.ZkbInvalidArgument):
        dg.qap_instance_evaluation(cs, dg.FIELD_BY_NAME["bn254_fr"], 5)
