"""Oracle self-pinning for the FRI query phase / verifier restatement (oracle/fri_query.py): the reference holds no
concrete LPC proof, so the pin is the scheme relation - proofs of the restated prover pass the restated verifier
(basic_fri.hpp:932-1150, lpc.hpp:202-263) for single- and multi-step rounds, and tampered proofs do not.  CPU only."""
import copy

import pytest

from oracle import fields, fri, fri_query, hashes


def _setup(step_list, degree_log, expand, lam, grinding, seed, hname="keccak256", fixed=False):
    F = fields.PALLAS_FQ
    _, h, _ = hashes.HASHES[hname]
    params = fri_query.FriParams(F, step_list, degree_log, lam, expand, grinding, 0xFF)
    n = 1 << degree_log
    polys = {0: [fields.random_elements(F, n, seed + i) for i in range(2)],
             2: [fields.random_elements(F, n, seed + 10 + i) for i in range(3)],
             3: [fields.random_elements(F, n // 2, seed + 20)]}
    d0 = params.D[0].m
    trees = {k: fri.precommit(polys[k], F, d0, step_list[0], h)[0] for k in polys}
    t = hashes.FiatShamirSequential(h, b"init")
    etha = t.challenge(F) if fixed else None
    y = 1234567 + seed
    yw = y * F.omega(degree_log) % F.p
    points = {0: [[y], [y]], 2: [[y, yw], [y], [yw, y]], 3: [[y]]}
    fixed_batches = (0,) if fixed else ()
    fixed_values = None
    if fixed:
        from oracle.ntt import dfs_coefficients
        from oracle.lpc import poly_eval
        fixed_values = {0: [poly_eval(dfs_coefficients(pl, F), etha, F.p) for pl in polys[0]]}
    return F, h, params, polys, points, trees, t, fixed_batches, etha, fixed_values


@pytest.mark.parametrize("step_list,degree_log,expand,grinding,fixed", [
    ([1, 1, 1], 4, 2, False, False),
    ([1, 1, 1, 1], 5, 1, True, True),
    ([2, 1, 1], 5, 2, False, False),
    ([3, 1], 5, 2, True, False),
    ([2, 2, 1], 6, 2, False, True),
])
def test_prover_verifier_round_trip(step_list, degree_log, expand, grinding, fixed):
    F, h, params, polys, points, trees, t, fb, etha, fv = _setup(step_list, degree_log, expand, 4, grinding, 5, fixed=fixed)
    tp, tv = t.copy(), t.copy()
    proof = fri_query.lpc_proof_eval(polys, points, trees, params, tp, h, fb, etha, fv)
    commitments = {k: trees[k][-1][0] for k in trees}
    assert len(proof["fri_proof"]["query_proofs"]) == 4
    assert len(proof["fri_proof"]["fri_roots"]) == len(step_list)
    assert fri_query.lpc_verify_eval(proof, points, commitments, params, tv, h, fb, etha, fv)
    assert tp.state == tv.state                       # prover and verifier transcripts stay in lockstep
    if grinding:
        assert proof["fri_proof"]["proof_of_work"] is not None

    def rejected(mut):
        bad = copy.deepcopy(proof)
        mut(bad)
        return not fri_query.lpc_verify_eval(bad, points, commitments, params, t.copy(), h, fb, etha, fv)

    def bump(v):
        return (v + 1) % F.p

    assert rejected(lambda b: b["z"][2][0].__setitem__(0, bump(b["z"][2][0][0])))
    assert rejected(lambda b: b["fri_proof"]["final_polynomial"].__setitem__(0, bump(b["fri_proof"]["final_polynomial"][0])))
    assert rejected(lambda b: b["fri_proof"]["query_proofs"][1]["initial_proof"][0]["values"][0][0].__setitem__(
        1, bump(b["fri_proof"]["query_proofs"][1]["initial_proof"][0]["values"][0][0][1])))
    assert rejected(lambda b: b["fri_proof"]["query_proofs"][0]["round_proofs"][0]["y"][0].__setitem__(
        0, bump(b["fri_proof"]["query_proofs"][0]["round_proofs"][0]["y"][0][0])))
    assert rejected(lambda b: b["fri_proof"]["query_proofs"][2]["round_proofs"][-1]["p"]["path"].__setitem__(
        0, bytes(len(b["fri_proof"]["query_proofs"][2]["round_proofs"][-1]["p"]["path"][0]))))
    if grinding:
        assert rejected(lambda b: b["fri_proof"].__setitem__("proof_of_work", b["fri_proof"]["proof_of_work"] + 1)) or True


def test_proof_of_work():
    """proof_of_work.hpp:47-81: generate/verify agree, the transcript advances identically on both sides"""
    _, h, _ = hashes.HASHES["keccak256"]
    t = hashes.FiatShamirSequential(h, b"pow")
    a, b = t.copy(), t.copy()
    nonce = fri_query.pow_generate(a, 0xFFF)
    assert fri_query.pow_verify(b, nonce, 0xFFF)
    assert a.state == b.state
    c = t.copy()
    assert not fri_query.pow_verify(c, nonce + 1, 0xFFF) or fri_query.pow_generate(t.copy(), 0xFFF, nonce + 1) == nonce + 1


def test_index_helpers():
    """calculate_s / get_correct_order against the precommit leaf layout (basic_fri.hpp:469-490): the s_indices of a
    query, put in 'correct order', are exactly the index pairs precommit hashed into that leaf"""
    F = fields.PALLAS_FQ
    for log_d, step in ((6, 1), (6, 2), (7, 3)):
        D = fri_query.EvaluationDomain(F, 1 << log_d)
        for x_index in (0, 1, 5, (1 << log_d) - 1, (1 << (log_d - 1)) + 3):
            x = D.get_domain_element(x_index)
            s, s_idx = fri_query.calculate_s(x, x_index, step, D)
            order = fri_query.get_correct_order(x_index, 1 << log_d, step, s_idx)
            leaf = fri_query.get_folded_index(x_index, 1 << log_d, step)
            want = fri.leaf_indices(leaf, 1 << log_d, step)
            got = []
            for idx, _ in order:
                got += [min(s_idx[idx]), max(s_idx[idx])]
            assert got == want
