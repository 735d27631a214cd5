"""Pins the oracle's Groth16 restatement (oracle/groth16.py) by the scheme's own relations: with the toxic waste
known, the proof elements must be (alpha + sum x_i A_i(t) + r delta) G1 and so on, and the QAP divisibility
A(t) B(t) - C(t) = H(t) Z(t) must hold at the secret point.  CPU only, small sizes."""
import pytest

from oracle import curves, fields, groth16


@pytest.mark.parametrize("kind", ["field", "binary"])
def test_groth16_oracle_satisfies_the_relations_in_the_exponent(kind):
    F, G1, G2 = fields.BN254_FR, curves.BN254_G1, curves.BN254_G2
    p = F.p
    if kind == "field":
        cs, primary, aux = groth16.example_with_field_input(F, 12, 3, seed=1)    # domain 16
    else:
        cs, primary, aux = groth16.example_with_binary_input(F, 11, 4, seed=2)   # domain 16
    t, alpha, beta, gamma, delta = fields.random_elements(F, 5, 11)
    pk = groth16.generator(cs, G1, G2, F, t, alpha, beta, gamma, delta)
    r, s = fields.random_elements(F, 2, 12)
    # QAP relation at t
    m, full, H = groth16.witness_map(pk.cs, primary, aux, F)
    x = [1] + full
    sc = pk.scalars
    At = sum(a * b for a, b in zip(x, sc["At"])) % p
    Bt = sum(a * b for a, b in zip(x, sc["Bt"])) % p
    Ct = sum(a * b for a, b in zip(x, sc["Ct"])) % p
    Hval = sum(h * pow(t, i, p) for i, h in enumerate(H)) % p
    assert (At * Bt - Ct) % p == Hval * sc["Zt"] % p
    assert m == 16 and H[m - 1] == 0 and H[m] == 0
    A, B, C = groth16.prove(pk, primary, aux, r, s, G1, G2, F)
    a, b, c = groth16.proof_in_the_exponent(pk, primary, aux, r, s, F)
    assert A == G1.mul(G1.gen, a)
    assert B == G2.mul(G2.gen, b)
    assert C == G1.mul(G1.gen, c)
