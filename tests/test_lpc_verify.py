"""The product-side verifier (crypto3_zk_b200/lpc_verify.py: lpc.hpp:202-263, basic_fri.hpp:932-1150) against proofs of the
oracle prover: accepts them, rejects tampered ones, and agrees with the oracle verifier.  CPU only."""
import copy

import pytest

from crypto3_zk_b200 import lpc_verify
from crypto3_zk_b200.lpc import FriParams
from crypto3_zk_b200.transcript import FiatShamirSequential
from oracle import fields, fri, fri_query, hashes, lpc, ntt

HID = {"keccak256": 0, "sha256": 1, "keccak512": 2}


@pytest.mark.parametrize("steps,degree_log,expand,grind,hname", [
    ([1, 1, 1], 4, 2, False, "keccak256"),
    ([2, 1, 1], 5, 2, True, "sha256"),
    ([3, 1], 5, 1, True, "keccak512"),
    ([2, 2, 1], 6, 2, False, "keccak256"),
])
def test_product_verifier_on_oracle_proofs(steps, degree_log, expand, grind, hname):
    F = fields.PALLAS_FQ
    p = F.p
    _, h, _ = hashes.HASHES[hname]
    n, d0, lam = 1 << degree_log, 1 << (degree_log + expand), 4
    polys = {0: [fields.random_elements(F, n, 1 + i) for i in range(2)], 1: [fields.random_elements(F, n, 9)],
             3: [fields.random_elements(F, n // 2, 20 + i) for i in range(2)]}
    params = fri_query.FriParams(F, steps, degree_log, lam, expand, grind, 0x1FF)
    trees = {k: fri.precommit(polys[k], F, d0, steps[0], h)[0] for k in polys}
    commitments = {k: trees[k][-1][0] for k in trees}
    t0 = hashes.FiatShamirSequential(h, b"\x03")
    etha = t0.challenge(F)
    fixed_values = {0: [lpc.poly_eval(ntt.dfs_coefficients(q, F), etha, p) for q in polys[0]]}
    y = 987654321
    points = {0: [[y], [y]], 1: [[y, y * F.omega(degree_log) % p]], 3: [[y], [y]]}
    proof = fri_query.lpc_proof_eval(polys, points, trees, params, t0.copy(), h, (0,), etha, fixed_values)
    fp = FriParams(steps, degree_log, lam, expand, grind, 0x1FF)

    def run(pr):
        tr = FiatShamirSequential(HID[hname], b"\x03")
        assert tr.challenge(p) == etha
        return lpc_verify.lpc_verify_eval(F.name, HID[hname], fp, pr, points, commitments, tr, (0,), etha, fixed_values), tr

    ok, tr = run(proof)
    assert ok
    # the same through the scheme object, used the way a verifier uses it: batch sizes, evaluation points, setup
    from crypto3_zk_b200.lpc import LpcCommitmentScheme
    vs = LpcCommitmentScheme(None, F.name, HID[hname], fp)
    for k in polys:
        vs.set_batch_size(k, len(polys[k]))
        for i, pts in enumerate(points[k]):
            for x in pts:
                vs.append_eval_point(k, x, poly=i)
    vs.mark_batch_as_fixed(0)
    tr2 = FiatShamirSequential(HID[hname], b"\x03")
    vs.setup(tr2, fixed_values)
    assert vs.verify_eval(proof, commitments, tr2) and tr2.state == tr.state
    tv = t0.copy()
    assert fri_query.lpc_verify_eval(proof, points, commitments, params, tv, h, (0,), etha, fixed_values)
    assert tr.state == tv.state                      # both verifiers leave the transcript in the same state
    for mutate in (
        lambda b: b["z"][1][0].__setitem__(1, (b["z"][1][0][1] + 1) % p),
        lambda b: b["fri_proof"]["final_polynomial"].__setitem__(0, (b["fri_proof"]["final_polynomial"][0] + 1) % p),
        lambda b: b["fri_proof"]["query_proofs"][3]["initial_proof"][3]["values"][1][0].__setitem__(
            0, (b["fri_proof"]["query_proofs"][3]["initial_proof"][3]["values"][1][0][0] + 1) % p),
        lambda b: b["fri_proof"]["query_proofs"][0]["round_proofs"][-1]["y"][0].__setitem__(
            1, (b["fri_proof"]["query_proofs"][0]["round_proofs"][-1]["y"][0][1] + 1) % p),
        lambda b: b["fri_proof"]["query_proofs"][1]["round_proofs"][0]["p"].__setitem__("index", 0 if b["fri_proof"]["query_proofs"][1]["round_proofs"][0]["p"]["index"] else 1),
        # malformed structures are rejected, never raised: missing query, missing round, missing batch, short path
        lambda b: b["fri_proof"]["query_proofs"].pop(),
        lambda b: b["fri_proof"]["query_proofs"][2]["round_proofs"].pop(),
        lambda b: b["z"].pop(3),
        lambda b: b["fri_proof"]["query_proofs"][0]["initial_proof"].pop(1),
        lambda b: b["fri_proof"]["query_proofs"][0]["initial_proof"][0]["p"]["path"].pop(),
        lambda b: b["fri_proof"]["fri_roots"].pop(),
    ):
        bad = copy.deepcopy(proof)
        mutate(bad)
        assert not run(bad)[0]


def test_merkle_validate_binds_index_and_depth():
    """a path that is consistent with the root but opens another leaf, or stops one level short (an inner
    node passed off as a leaf), must not validate."""
    h = hashes.HASHES["keccak256"][1]
    leaves = [bytes([i]) * 64 for i in range(8)]
    lvl = [h(x) for x in leaves]
    levels = [lvl]
    while len(lvl) > 1:
        lvl = [h(lvl[i] + lvl[i + 1]) for i in range(0, len(lvl), 2)]
        levels.append(lvl)
    root = levels[-1][0]

    def path(i):
        out = []
        for l in levels[:-1]:
            out.append(l[i ^ 1])
            i >>= 1
        return out
    good = {"index": 5, "path": path(5), "root": root}
    assert lpc_verify._merkle_validate(good, leaves[5], h, 5, 3)
    assert not lpc_verify._merkle_validate(good, leaves[5], h, 4, 3)            # not the queried leaf
    other = {"index": 4, "path": path(4), "root": root}
    assert not lpc_verify._merkle_validate(other, leaves[4], h, 5, 3)           # a valid opening of another leaf
    # inner node (h(l4) || h(l5)) presented as a 64-byte leaf with a path one level short
    inner = {"index": 2, "path": path(5)[1:], "root": root}
    assert lpc_verify._merkle_validate.__code__.co_argcount == 5
    assert not lpc_verify._merkle_validate(inner, levels[0][4] + levels[0][5], h, 2, 3)
