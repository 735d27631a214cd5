"""Builds and runs the C++ host-template test (tests/cpp/host_api_test.cpp) that mirrors the reference's
own Boost tests for the hot-path entities on top of the C ABI."""
import os
import subprocess

import pytest

from oracle import fields, fri, fri_query, hashes, lpc, ntt

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "tests", "cpp", "host_api_test")


def _build():
    from crypto3_zk_b200 import build
    build.build()
    src = os.path.join(ROOT, "tests", "cpp", "host_api_test.cpp")
    deps = [src, os.path.join(ROOT, "crypto3_zk_b200", "host", "zkb_crypto3.hpp"),
            os.path.join(ROOT, "crypto3_zk_b200", "host", "zkb_r1cs_gg_ppzksnark.hpp"),
            os.path.join(ROOT, "crypto3_zk_b200", "host", "zkb_placeholder.hpp"), os.path.join(ROOT, "include", "zkb200.h")]
    if not os.path.exists(EXE) or any(os.path.getmtime(d) > os.path.getmtime(EXE) for d in deps):
        lib = os.path.join(ROOT, "crypto3_zk_b200")
        subprocess.check_call(["g++", "-O2", "-std=c++17", src, "-o", EXE, "-L" + lib, "-lzkb200", "-Wl,-rpath," + lib])
    return EXE


def test_host_templates_compile_and_link():
    exe = _build()
    out = subprocess.run([exe, "compile-only"], stdout=subprocess.PIPE, text=True, check=True).stdout
    assert "compiled" in out
    # the C++ curve_element_serializer / proof blob against the Python layer's wire format (host only)
    from crypto3_zk_b200 import marshalling as m
    from oracle import curves
    g1, g2 = curves.BLS12_381_G1, curves.BLS12_381_G2
    lines = _marshal_lines(out)
    assert lines["g1gen"] == m.g1_to_bytes(g1.gen).hex() and lines["g2gen"] == m.g2_to_bytes(g2.gen).hex()
    assert lines["proof"] == m.proof_to_bytes((g1.gen, g2.neg(g2.gen), g1.neg(g1.gen))).hex()
    assert lines["pi"] == m.primary_input_to_bytes([1, m.R - 1, 12345, 0]).hex()


def _marshal_lines(out):
    return {l.split()[1]: l.split()[2] for l in out.splitlines() if l.startswith("MARSHAL ")}


@pytest.mark.gpu
def test_host_templates_on_gpu():
    exe = _build()
    r = subprocess.run([exe], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout
    root = [l.split()[1] for l in r.stdout.splitlines() if l.startswith("ROOT ")][0]
    F = fields.PALLAS_FP
    polys = [[(p + 1) * 1000 + i for i in range(16)] for p in range(3)]
    levels, _ = fri.precommit(polys, F, 64, 2, hashes.keccak256)
    assert root == levels[-1][0].hex()
    # lpc_commitment_scheme of the host templates: same proof as the oracle prover for the same inputs
    got = [l.split() for l in r.stdout.splitlines() if l.startswith("PROOF ")][0]
    want_digest, want_len, want_state = _oracle_lpc_proof_digest()
    assert (got[1], int(got[2])) == (want_digest, want_len)
    assert [l.split()[1] for l in r.stdout.splitlines() if l.startswith("TRANSCRIPT ")][0] == want_state
    # compressed encodings of device-computed multiples of the generators: same bytes as the Python layer
    # produces for the oracle's scalar multiples
    from crypto3_zk_b200 import marshalling as m
    from oracle import curves
    lines = _marshal_lines(r.stdout)
    for k in (2, 3, 12345678901, 0xFFFFFFFFFFFFFFFF):
        assert lines["g1x%d" % k] == m.g1_to_bytes(curves.BLS12_381_G1.mul(curves.BLS12_381_G1.gen, k)).hex()
        assert lines["g2x%d" % k] == m.g2_to_bytes(curves.BLS12_381_G2.mul(curves.BLS12_381_G2.gen, k)).hex()


def _oracle_lpc_proof_digest():
    """tests/cpp/host_api_test.cpp:lpc_scheme_test restated with the oracle prover; same canonical dump"""
    F, h = fields.PALLAS_FP, hashes.keccak256
    p = F.p

    def poly(n, seed):
        return [(seed * 1000003 + i * i * 7 + i + 1) % p for i in range(n)]

    polys = {0: [poly(32, 1 + i) for i in range(2)], 1: [poly(32, 10 + i) for i in range(3)], 4: [poly(16, 20)]}
    params = fri_query.FriParams(F, [2, 1, 1], 5, 5, 2, True, 0x3FF)
    tr = hashes.FiatShamirSequential(h, bytes([7]))
    etha = tr.challenge(F)
    fixed_values = {0: [lpc.poly_eval(ntt.dfs_coefficients(q, F), etha, p) for q in polys[0]]}
    trees = {k: fri.precommit(polys[k], F, 128, 2, h)[0] for k in polys}
    y = 123456789
    yw = y * F.omega(5) % p
    points = {0: [[y], [y]], 1: [[y], [y, yw], [yw]], 4: [[y]]}
    proof = fri_query.lpc_proof_eval(polys, points, trees, params, tr, h, (0,), etha, fixed_values)
    d = bytearray()

    def val(v):
        d.extend(int(v).to_bytes(32, "big"))

    def mp(q):
        d.extend(int(q["index"]).to_bytes(8, "big"))
        for s in q["path"]:
            d.extend(s)
        d.extend(q["root"])

    for k in sorted(proof["z"]):
        for pz in proof["z"][k]:
            for v in pz:
                val(v)
    f = proof["fri_proof"]
    for r_ in f["fri_roots"]:
        d.extend(r_)
    for v in f["final_polynomial"]:
        val(v)
    d.extend(int(f["proof_of_work"]).to_bytes(4, "big"))
    for q in f["query_proofs"]:
        for k in sorted(q["initial_proof"]):
            ip = q["initial_proof"][k]
            for pv in ip["values"]:
                for pr in pv:
                    val(pr[0])
                    val(pr[1])
            mp(ip["p"])
        for rp in q["round_proofs"]:
            for pr in rp["y"]:
                val(pr[0])
                val(pr[1])
            mp(rp["p"])
    return hashes.keccak256(bytes(d)).hex(), len(d), tr.state.hex()


def _g(pt, deg=1):
    if pt is None:
        return "inf"
    if deg == 1:
        return "%x %x" % (pt[0], pt[1])
    return "%x %x %x %x" % (pt[0][0], pt[0][1], pt[1][0], pt[1][1])


@pytest.mark.gpu
@pytest.mark.parametrize("curve,kind,nc,ni", [("bn254", "field", 28, 3), ("bls12_381", "binary", 59, 4)])
def test_cpp_groth16_prover_vs_oracle(tmp_path, curve, kind, nc, ni):
    """C++ r1cs_gg_ppzksnark_prover<Curve>::process and r1cs_to_qap<F>::witness_map (host/zkb_r1cs_gg_ppzksnark.hpp, the
    counterparts of prover.hpp:73-158 and r1cs_to_qap.hpp:219-325) on a key made by the oracle's generator: same H
    coefficients (with and without the d1/d2/d3 patch) and the same proof (affine) as the oracle prover."""
    from oracle import curves, groth16
    F, G1, G2 = ((fields.BN254_FR, curves.BN254_G1, curves.BN254_G2) if curve == "bn254" else
                 (fields.BLS12_381_FR, curves.BLS12_381_G1, curves.BLS12_381_G2))
    make = groth16.example_with_field_input if kind == "field" else groth16.example_with_binary_input
    cs, primary, aux = make(F, nc, ni, seed=11)
    t, alpha, beta, gamma, delta = fields.random_elements(F, 5, 41)
    pk = groth16.generator(cs, G1, G2, F, t, alpha, beta, gamma, delta)
    r, s = fields.random_elements(F, 2, 42)
    cs = pk.cs       # the generator may have swapped A and B (swap_AB_if_beneficial, generator.hpp:104-106)
    lines = [curve, "%d %d %d" % (cs.num_inputs, cs.num_aux, cs.num_constraints)]
    for con in cs.constraints:
        for side in con:
            lines.append(" ".join([str(len(side))] + ["%d %x" % (i, co % F.p) for i, co in side]))
    lines += [_g(pk.alpha_g1), _g(pk.beta_g1), _g(pk.beta_g2, 2), _g(pk.delta_g1), _g(pk.delta_g2, 2)]
    lines.append(str(len(pk.A_query)))
    lines += [_g(q) for q in pk.A_query]
    lines.append("%d %d" % (len(pk.B_indices), len(pk.A_query)))
    lines += ["%d %s %s" % (i, _g(g2, 2), _g(g1)) for i, g2, g1 in zip(pk.B_indices, pk.B_g2, pk.B_g1)]
    lines.append(str(len(pk.H_query)))
    lines += [_g(q) for q in pk.H_query]
    lines.append(str(len(pk.L_query)))
    lines += [_g(q) for q in pk.L_query]
    lines.append(" ".join("%x" % v for v in primary))
    lines.append(" ".join("%x" % v for v in aux))
    lines.append("%x %x" % (r, s))
    path = tmp_path / "groth16_key.txt"
    path.write_text("\n".join(lines) + "\n")
    exe = _build()
    res = subprocess.run([exe, "groth16", str(path)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert res.returncode == 0 and "ALL OK" in res.stdout, res.stdout[-2000:]
    out = {l.split()[0]: l.split()[1:] for l in res.stdout.splitlines() if l.startswith("G16")}
    m, full, H = groth16.witness_map(pk.cs, primary, aux, F)
    assert [int(v, 16) for v in out["G16H"]] == H[:m + 1]
    _, _, HD = groth16.witness_map(pk.cs, primary, aux, F, 3, 5, 7)
    assert [int(v, 16) for v in out["G16HD"]] == HD[:m + 1]
    want = groth16.prove(pk, primary, aux, r, s, G1, G2, F)
    got = [[int(v, 16) for v in part.split()] for part in " ".join(out["G16PROOF"]).split(" | ")]
    assert got[0] == list(want[0]) and got[1] == [want[1][0][0], want[1][0][1], want[1][1][0], want[1][1][1]] and got[2] == list(want[2])


@pytest.mark.gpu
@pytest.mark.parametrize("log_n,triples,mqc,lookup", [(4, 1, 0, False), (6, 2, 4, False), (6, 1, 5, True), (7, 1, 0, True)])
def test_cpp_placeholder_prover_vs_python_driver(tmp_path, log_n, triples, mqc, lookup):
    """C++ placeholder_prover<F, Hash, Hash>::preprocess / process (host/zkb_placeholder.hpp, after placeholder/prover.hpp:
    133-217, permutation_argument.hpp:95-215, gates_argument.hpp:133-217) on the chain circuit: the same commitments,
    challenge, opened values, FRI roots and transcript state as crypto3_zk_b200.placeholder.placeholder_prove, which
    tests/test_gpu_placeholder.py checks against the oracle and the verifier's identity."""
    import numpy as np
    from crypto3_zk_b200 import Context, placeholder as P, workloads as W
    from crypto3_zk_b200.lpc import FriParams
    from crypto3_zk_b200.transcript import FiatShamirSequential
    F = fields.PALLAS_FP
    ctx = Context(0)
    try:
        circuit, witness, public = W.placeholder_chain_circuit(ctx, F.name, log_n, triples=triples, seed=log_n, max_quotient_chunks=mqc,
                                                                lookup=lookup)

        def dump(t):
            a = t.cpu().numpy().view(np.uint32).reshape(-1, 8)
            return " ".join("%x" % v for v in fields.u32_array_to_ints(a))

        lam, expand = 4, 3
        lines = ["%d %d %d %d %d %d %d" % (log_n, triples, circuit.usable_rows, mqc, lam, expand, int(lookup))]
        lines += [dump(witness), dump(public)] + ([dump(circuit.constants)] if lookup else [])
        lines += [dump(circuit.selectors), dump(circuit.s_id), dump(circuit.s_sigma),
                  dump(circuit.q_last), dump(circuit.q_blind), dump(circuit.lagrange_0)]
        path = tmp_path / "placeholder.txt"
        path.write_text("\n".join(lines) + "\n")
        tr = FiatShamirSequential(0, b"placeholder-test")
        res = P.placeholder_prove(ctx, circuit, 0, FriParams.with_max_step_one(log_n, lam, expand), witness, public, tr)
        state = tr.state
    finally:
        ctx.close()
    exe = _build()
    r = subprocess.run([exe, "placeholder", str(path)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    assert r.returncode == 0 and "ALL OK" in r.stdout, r.stdout[-2000:]
    out = {l.split()[1]: l.split()[2:] for l in r.stdout.splitlines() if l.startswith("PLH ")}
    assert out["fixed"][0] == res["commitments"][0].hex()
    for k in (1, 2, 3) + ((4,) if lookup else ()):
        assert out["root%d" % k][0] == res["commitments"][k].hex(), k
    assert int(out["y"][0], 16) == res["challenge"]
    assert [int(v) for v in out["chunks"]] == [res["quotient_chunks"], res["log_d"]]
    z = res["eval_proof"]["z"]
    for k in (0, 1, 2, 3) + ((4,) if lookup else ()):
        got = [[int(v, 16) for v in part.split()] for part in " ".join(out["z%d" % k]).split("|")[1:]]
        assert got == [list(v) for v in z[k]], k
    assert out["fri"] == [rt.hex() for rt in res["eval_proof"]["fri"]["roots"]]
    assert out["transcript"][0] == bytes(state).hex()
