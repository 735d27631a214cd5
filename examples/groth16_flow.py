#!/usr/bin/env python3
"""Groth16 on BLS12-381 between the reference's byte formats (r1cs_gg_ppzksnark: generator.hpp:83-235, prover.hpp:73-158,
marshalling.hpp): generate the key of a multiplication chain on the GPU, write it as the reference's proving-key blob, read
it back, prove, and write the 192-byte proof.  Self-check: g_A and g_B equal their discrete logs (known from the toxic waste)
times the generators.  The same steps are tests/test_gpu_flows.py::test_groth16_generator_vs_oracle and the wire-format leg
of test_groth16_prove_vs_oracle, checked there against the CPU oracle."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from crypto3_zk_b200 import Context, groth16, marshalling  # noqa: E402
from crypto3_zk_b200.api import _affine_from_limbs, _int_rows  # noqa: E402
from crypto3_zk_b200.fields import CURVE_BY_NAME, FIELD_BY_NAME, coord_limbs  # noqa: E402


def chain(p, log_m, num_inputs, rng):
    """x[i+2] = x[i] * x[i+1] (r1cs_examples.hpp:77-146 in spirit), sized so that the domain is exactly 2^log_m"""
    nc = (1 << log_m) - num_inputs - 1
    x = [int.from_bytes(rng.bytes(32), "little") % p for _ in range(2)]
    cons = []
    for i in range(nc):
        cons.append(([(i + 1, 1)], [(i + 2, 1)], [(i + 3, 1)]))
        x.append(x[i] * x[i + 1] % p)
    return groth16.R1csConstraintSystem(num_inputs, len(x) - num_inputs, cons), x[:num_inputs], x[num_inputs:]


def main(log_m=10):
    ctx = Context(0)
    import torch
    torch.zeros(1, device="cuda:0")        # torch's own CUDA start-up (seconds on a fresh box) stays out of the timings
    G1, G2, F = CURVE_BY_NAME["bls12_381_g1"], CURVE_BY_NAME["bls12_381_g2"], FIELD_BY_NAME["bls12_381_fr"]
    p = F.p
    rng = np.random.Generator(np.random.PCG64(3))
    rnd = lambda: int.from_bytes(rng.bytes(32), "little") % p
    cs, primary, aux = chain(p, log_m, 2, rng)
    t, alpha, beta, gamma, delta, r, s = (rnd() for _ in range(7))
    t0 = time.perf_counter()
    key, vk = groth16.generator(ctx, G1.name, G2.name, cs, t, alpha, beta, gamma, delta)
    t1 = time.perf_counter()
    blob = marshalling.proving_key_to_bytes(key)
    t1b = time.perf_counter()
    key2 = marshalling.proving_key_from_bytes(blob, ctx=ctx)     # query vectors decompressed on the device
    t2 = time.perf_counter()
    pk = groth16.proving_key_from_dict(ctx, G1.name, G2.name, key2)
    proof = groth16.prove(ctx, pk, primary, aux, r, s)
    t3 = time.perf_counter()
    wire = marshalling.proof_to_bytes(proof)
    assert marshalling.proof_from_bytes(wire) == proof
    # the proof in the exponent: A = alpha + sum x_i A_i(t) + r delta, B = beta + sum x_i B_i(t) + s delta
    At, Bt, _, _, _, _ = groth16.qap_instance_evaluation(pk.cs, F, t)
    x = [1] + primary + aux
    a = (alpha + sum(xi * ai for xi, ai in zip(x, At)) + r * delta) % p
    b = (beta + sum(xi * bi for xi, bi in zip(x, Bt)) + s * delta) % p
    ea = ctx.batch_exp(G1.name, (G1.gen_x, G1.gen_y), _int_rows([a]))
    eb = ctx.batch_exp(G2.name, (G2.gen_x, G2.gen_y), _int_rows([b]))
    assert proof[0] == _affine_from_limbs(np.asarray(ea)[0].reshape(-1), coord_limbs(G1), 1), "g_A"
    assert proof[1] == _affine_from_limbs(np.asarray(eb)[0].reshape(-1), coord_limbs(G2), 2), "g_B"
    print("Groth16, domain 2^%d: generator %.1f ms, key blob %d bytes (write %.1f ms on the host, read %.1f ms with the points "
          "decompressed on the device), key + proof %.1f ms, proof %s...; g_A, g_B match their discrete logs" %
          (log_m, (t1 - t0) * 1e3, len(blob), (t1b - t1) * 1e3, (t2 - t1b) * 1e3, (t3 - t2) * 1e3, wire[:8].hex()))
    ctx.close()


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 10)
