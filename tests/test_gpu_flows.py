"""GPU parity of the callers of the hot path through the C ABI:
FRI commit phase, eval_polys, combined Q, and the lpc_commitment_scheme flow, bit-exact against the oracle."""
import numpy as np
import pytest

from oracle import fields, fri, hashes, lpc, ntt

pytestmark = pytest.mark.gpu

HASHES = {0: hashes.keccak256, 1: hashes.sha256, 2: hashes.keccak512}


@pytest.fixture(scope="module")
def ctx():
    from crypto3_zk_b200 import Context
    c = Context(0)
    yield c
    c.close()


def to_arr(vals, limbs=8):
    return fields.ints_to_u32_array(vals, limbs)


def from_arr(a):
    return fields.u32_array_to_ints(np.asarray(a).reshape(-1, a.shape[-1]))


def dev(a):
    import torch
    return torch.from_numpy(np.ascontiguousarray(a).view(np.int32)).cuda()


def host(t):
    return t.cpu().numpy().view(np.uint32)


class OracleTranscript:
    """adapts oracle.hashes.FiatShamirSequential to the product's transcript interface"""

    def __init__(self, h, F, init=b"\x00"):
        self.t, self.F = hashes.FiatShamirSequential(h, init), F
        self.hash_id = [k for k, v in HASHES.items() if v is h][0]

    def __call__(self, data):
        self.t.absorb(data)

    def challenge(self, modulus):
        assert modulus == self.F.p
        return self.t.challenge(self.F)

    def int_challenge(self, bits=32):
        return self.t.int_challenge(bits)

    @property
    def state(self):
        return self.t.state


@pytest.mark.parametrize("hid", [0, 1, 2], ids=["keccak256", "sha256", "keccak512"])
@pytest.mark.parametrize("F,log_n,steps", [(fields.PALLAS_FQ, 6, [1, 1, 1]), (fields.PALLAS_FP, 8, [2, 1, 3]),
                                           (fields.BLS12_381_FR, 7, [3, 2, 1, 1]), (fields.BN254_FR, 5, [1])],
                         ids=lambda v: getattr(v, "name", str(v)))
def test_fri_commit_phase_vs_oracle(ctx, hid, F, log_n, steps):
    """zk::algorithms::proof_eval<FRI> commit phase (basic_fri.hpp:706-737)."""
    h = HASHES[hid]
    f = fields.random_elements(F, 1 << log_n, 11 + log_n)
    want = fri.commit_phase(f, F, log_n, steps, h, hashes.FiatShamirSequential(h, b"init"))
    tr = OracleTranscript(h, F, b"init")

    def challenge(rnd, root, count):
        tr(root)
        return [tr.challenge(F.p) for _ in range(count)]

    for data in (to_arr(f), dev(to_arr(f))):
        tr = OracleTranscript(h, F, b"init")
        got = ctx.fri_commit_phase(F.name, hid, data, log_n, steps, challenge, keep_trees=True, keep_fs=True)
        assert got["roots"] == want["roots"]
        assert got["alphas"] == want["alphas"]
        assert got["final_polynomial"] == want["final_polynomial"]
        assert from_arr(host(got["fs"])) == [v for fs in want["fs"][1:] for v in fs]
        for tree, levels in zip(got["trees"], want["levels"]):
            assert tree.root() == levels[-1][0]
            idx = min(1, tree.leaves - 1)
            assert tree.path(idx) == fri.merkle_proof(levels, idx)


def test_fri_commit_phase_errors(ctx):
    from crypto3_zk_b200 import capi
    F = fields.PALLAS_FQ
    f = to_arr(fields.random_elements(F, 16, 1))
    with pytest.raises(capi.ZkbInvalidArgument):
        ctx.fri_commit_phase(F.name, 0, f, 4, [3, 2], lambda r, root, c: [1] * c)
    with pytest.raises(capi.ZkbInvalidArgument):
        ctx.fri_commit_phase(F.name, 0, f, 4, [], lambda r, root, c: [1] * c)
    with pytest.raises(ZeroDivisionError):   # an exception in the caller's transcript surfaces unchanged
        ctx.fri_commit_phase(F.name, 0, f, 4, [1, 1], lambda r, root, c: [1 // 0])


def test_fri_commit_phase_large_matches_coefficient_folds(ctx):
    """Size-independent identity at a Placeholder-sized domain (2^20 rows, blow-up 4): the final polynomial equals
    the coefficient folds f_even + alpha f_odd of the committed polynomial with the same challenges."""
    import torch
    from crypto3_zk_b200.transcript import FiatShamirSequential
    F = fields.PALLAS_FQ
    log_deg, log_n = 12, 22
    co = fields.random_elements(F, 1 << log_deg, 77)
    x = torch.zeros((1, 1 << log_n, 8), dtype=torch.int32, device="cuda")
    x[0, :1 << log_deg] = dev(to_arr(co))
    ctx.ntt(F.name, x, log_n)
    tr = FiatShamirSequential(0)
    steps = [1] * (log_deg - 1)
    got = ctx.fri_commit_phase(F.name, 0, x[0], log_n, steps, lambda r, root, c: (tr(root), tr.challenges(F.p, c))[1])
    c = co
    for a in got["alphas"]:
        c = fri.fold_polynomial_coeffs(c, a, F.p)
    fin = got["final_polynomial"]
    assert len(fin) == 1 << (log_n - sum(steps))
    assert fin[:len(c)] == c and not any(fin[len(c):])


@pytest.mark.parametrize("F", [fields.PALLAS_FQ, fields.BLS12_381_FR], ids=lambda f: f.name)
@pytest.mark.parametrize("n,batch,npts", [(1, 1, 1), (5, 2, 1), (256, 3, 2), (1000, 2, 5), (5000, 4, 9)])
def test_poly_evaluate_vs_oracle(ctx, F, n, batch, npts):
    polys = [fields.random_elements(F, n, 100 + b) for b in range(batch)]
    pts = fields.random_elements(F, npts, 9) if npts > 1 else [0]
    want = [[lpc.poly_eval(c, x, F.p) for x in pts] for c in polys]
    a = to_arr([v for c in polys for v in c]).reshape(batch, n, 8)
    assert ctx.poly_evaluate(F.name, a, n, pts) == want
    assert ctx.poly_evaluate(F.name, dev(a), n, pts) == want


@pytest.mark.parametrize("n,batch,npts", [(1, 1, 1), (5, 2, 3), (255, 2, 4), (256, 1, 5), (1000, 3, 2), (5000, 2, 9), (70000, 2, 6)])
def test_poly_evaluate_pm_vs_oracle(ctx, n, batch, npts):
    """values at z and -z from one pass (query phase openings, basic_fri.hpp:819-834)"""
    F = fields.PALLAS_FP
    polys = [fields.random_elements(F, n, 31 + b) for b in range(batch)]
    pts = fields.random_elements(F, npts, 5)
    got = ctx.poly_evaluate_pm(F.name, dev(to_arr([v for c in polys for v in c]).reshape(batch, n, 8)), n, pts)
    want = [[(lpc.poly_eval(c, z, F.p), lpc.poly_eval(c, (F.p - z) % F.p, F.p)) for z in pts] for c in polys]
    assert got == want


def test_poly_evaluate_dfs_and_large(ctx):
    """polynomial_dfs::evaluate = coefficients() then the value; 2^20-sized columns against Horner on the CPU."""
    F = fields.PALLAS_FQ
    vals = [fields.random_elements(F, 64, 5 + b) for b in range(3)]
    pts = fields.random_elements(F, 3, 6)
    a = to_arr([v for c in vals for v in c]).reshape(3, 64, 8)
    assert ctx.poly_evaluate(F.name, dev(a), 64, pts, dfs=True) == [[ntt.dfs_evaluate(v, F, x) for x in pts] for v in vals]
    n = 1 << 20
    rng = np.random.Generator(np.random.PCG64(5))
    big = rng.integers(0, 1 << 32, size=(2, n, 8), dtype=np.uint64).astype(np.uint32)
    big[..., 7] &= 0x0FFFFFFF
    x = pts[0]
    got = ctx.poly_evaluate(F.name, dev(big), n, [x, 1])
    for b in range(2):
        co = fields.u32_array_to_ints(big[b])
        assert got[b][0] == lpc.poly_eval(co, x, F.p)
        assert got[b][1] == sum(co) % F.p


@pytest.mark.parametrize("F", [fields.PALLAS_FP, fields.BN254_FR], ids=lambda f: f.name)
@pytest.mark.parametrize("n", [1, 2, 7, 8, 9, 2047, 2048, 2049, 5000, 1 << 15])
def test_poly_div_linear_vs_oracle(ctx, F, n):
    c = fields.random_elements(F, n, 200 + n)
    z = fields.random_elements(F, 1, 3)[0]
    q, rem = ctx.poly_div_linear(F.name, dev(to_arr(c)), n, z)
    assert rem == lpc.poly_eval(c, z, F.p)
    assert from_arr(host(q)) == lpc.poly_div_linear(c, z, F.p) + [0]


def test_poly_lincomb_vs_oracle(ctx):
    F = fields.PALLAS_FQ
    p, n, batch = F.p, 300, 5
    polys = [fields.random_elements(F, n, 300 + b) for b in range(batch)]
    sc = fields.random_elements(F, batch, 7)
    sc[2] = 0
    const = 123456789
    a = dev(to_arr([v for c in polys for v in c]).reshape(batch, n, 8))
    want = [sum(s * c[i] for s, c in zip(sc, polys)) % p for i in range(n)]
    want0 = list(want)
    want0[0] = (want0[0] - const) % p
    out = ctx.poly_lincomb(F.name, a, n, sc, constant=const)
    assert from_arr(host(out)) == want0
    ctx.poly_lincomb(F.name, a, n, sc, out=out, accumulate=True)
    assert from_arr(host(out)) == [(x + y) % p for x, y in zip(want0, want)]


@pytest.mark.parametrize("F,hid", [(fields.PALLAS_FQ, 0), (fields.BLS12_381_FR, 1)], ids=["pallas-keccak", "bls-sha256"])
def test_lpc_scheme_flow_vs_oracle(ctx, F, hid):
    """lpc_commitment_scheme: commit per batch (lpc.hpp:101-106), eval_polys (batched_commitment.hpp:176-190),
    combined Q (lpc.hpp:126-181) and the FRI commit phase (basic_fri.hpp:706-737), in the reference's transcript order."""
    from crypto3_zk_b200.lpc import FriParams, LpcCommitmentScheme
    h, p = HASHES[hid], F.p
    degree_log, expand, steps = 5, 2, [1, 2, 1]
    n, log_d0 = 1 << degree_log, degree_log + expand
    polys = {0: [fields.random_elements(F, n, 1 + i) for i in range(2)],
             1: [fields.random_elements(F, n, 10 + i) for i in range(3)],
             3: [fields.random_elements(F, n, 20)]}
    y = fields.random_elements(F, 1, 99)[0]
    yw = y * F.omega(degree_log) % p
    points = {0: [[y], [y]], 1: [[y], [y, yw], [yw]], 3: [[y]]}
    # ---- oracle, in the reference's order: setup (etha), commits, proof_eval
    tr = hashes.FiatShamirSequential(h, b"\x05")
    etha = tr.challenge(F)
    roots = {k: fri.lpc_commit(polys[k], F, degree_log, expand, steps[0], h) for k in polys}
    z = lpc.eval_polys(polys, points, F)
    for k in sorted(roots):
        tr.absorb(roots[k])
    theta = tr.challenge(F)
    fixed_values = {0: [lpc.poly_eval(ntt.dfs_coefficients(q, F), etha, p) for q in polys[0]]}
    q_normal, q_dfs = lpc.combined_q(polys, points, z, theta, F, fixed_batches=(0,), etha=etha, fixed_values=fixed_values)
    q_d0 = ntt.dfs_resize(q_dfs, F, 1 << log_d0)
    want = fri.commit_phase(q_d0, F, log_d0, steps, h, tr)
    # ---- device
    scheme = LpcCommitmentScheme(ctx, F.name, hid, FriParams(steps, degree_log, 40, expand))
    t2 = OracleTranscript(h, F, b"\x05")
    for k in polys:
        scheme.append_to_batch(k, dev(to_arr([v for c in polys[k] for v in c]).reshape(len(polys[k]), n, 8)))
    scheme.mark_batch_as_fixed(0)
    scheme.setup(t2, fixed_values)
    assert {k: scheme.commit(k) for k in polys} == roots
    for k in points:
        for i, pts in enumerate(points[k]):
            for x in pts:
                scheme.append_eval_point(k, x, poly=i)
    got = scheme.proof_eval(t2, keep_fs=True)
    assert got["z"] == z
    assert got["theta"] == theta
    assert got["remainders"] == [0] * len(got["remainders"]) and len(got["remainders"]) == 3
    assert from_arr(host(got["combined_Q_normal"])) == q_normal + [0] * (n - len(q_normal))
    assert from_arr(host(got["combined_Q"])) == q_d0
    assert got["fri"]["roots"] == want["roots"]
    assert got["fri"]["alphas"] == want["alphas"]
    assert got["fri"]["final_polynomial"] == want["final_polynomial"]


# ------------------------------------------------------------------------------------------ query phase
@pytest.mark.parametrize("hid", [0, 1, 2], ids=["keccak256", "sha256", "keccak512"])
@pytest.mark.parametrize("mask", [0xFF, 0xFFF, 0x3FFFFF])
def test_pow_grind_vs_oracle(ctx, hid, mask):
    """proof_of_work<Hash, uint32>::generate (proof_of_work.hpp:47-68): smallest passing nonce, any start"""
    from oracle import fri_query
    h = HASHES[hid]
    t = hashes.FiatShamirSequential(h, b"grind-%d" % mask)
    nonce = ctx.pow_grind(hid, t.state, mask)
    assert fri_query.pow_verify(t.copy(), nonce, mask)
    if mask <= 0xFFF:
        assert nonce == fri_query.pow_generate(t.copy(), mask)
        nxt = ctx.pow_grind(hid, t.state, mask, start=nonce + 1)
        assert nxt > nonce and nxt == fri_query.pow_generate(t.copy(), mask, nonce + 1)
    else:   # the oracle walk would take minutes: check minimality on a window below the hit instead
        lo = max(0, nonce - 300)
        for cand in range(lo, nonce):
            assert not fri_query.pow_verify(t.copy(), cand, mask)


@pytest.mark.parametrize("log_in,log_out", [(4, 4), (5, 8), (10, 13), (16, 19)])
def test_lde_with_coefficients(ctx, log_in, log_out):
    """zkb_lde_with_coefficients: the resize and the coefficient form it passes through (= inverse_fft of the input)"""
    import torch
    F = fields.BLS12_381_FR
    x = dev(to_arr(fields.random_elements(F, 3 << log_in, 17)).reshape(3, 1 << log_in, 8))
    co = torch.empty_like(x)
    ext = ctx.lde(F.name, x, log_in, log_out, coefficients_out=co)
    assert torch.equal(ext, ctx.lde(F.name, x, log_in, log_out))
    assert torch.equal(co, ctx.ntt(F.name, x.clone(), log_in, inverse=True))
    if log_in <= 10:
        want = [v for b in range(3) for v in ntt.dfs_coefficients(from_arr(host(x[b])), F)]
        assert from_arr(host(co)) == want


def test_query_phase_entry_points_edge_cases(ctx):
    """argument checking and degenerate sizes of the 8(f)-2 entry points"""
    import ctypes
    from crypto3_zk_b200 import capi
    F = fields.PALLAS_FQ
    st = hashes.keccak256(b"x")
    assert ctx.pow_grind(0, st, 0) == 0 and ctx.pow_grind(0, st, 0, start=77) == 77      # empty mask: the start itself
    with pytest.raises(capi.ZkbInvalidArgument):
        ctx.pow_grind(0, st[:31], 0xFF)
    with pytest.raises(capi.ZkbInvalidArgument):
        ctx.pow_grind(7, st, 0xFF)
    nonce = ctypes.c_uint32(0)
    buf = (ctypes.c_uint8 * 32).from_buffer_copy(st)
    assert capi.lib().zkb_pow_grind(ctx._h, 0, None, 0, 1, ctypes.byref(nonce), None) == capi.ERR_INVALID_ARGUMENT
    assert capi.lib().zkb_pow_grind(ctx._h, 0, buf, 0, 1, None, None) == capi.ERR_INVALID_ARGUMENT
    # a single polynomial of two evaluations, fri_step 1: one leaf, depth 0
    one = ctx.merkle_commit(F.name, 0, dev(to_arr([5, 9]).reshape(1, 2, 8)), 1, 1, keep_tree=True)
    assert one.leaves == 1 and one.paths([0, 0]) == [[], []]
    assert one.root() == hashes.keccak256((5).to_bytes(32, "big") + (9).to_bytes(32, "big"))
    # constants and tiny polynomials through the paired evaluation
    x = dev(to_arr([3, 4, 5]).reshape(3, 1, 8))
    assert ctx.poly_evaluate_pm(F.name, x, 1, [0, 1, F.p - 1]) == [[(c, c)] * 3 for c in (3, 4, 5)]
    y = dev(to_arr([2, 7]).reshape(1, 2, 8))
    assert ctx.poly_evaluate_pm(F.name, y, 2, [10]) == [[(72, (2 - 70) % F.p)]]
    with pytest.raises(ValueError):
        ctx.poly_evaluate_pm(F.name, to_arr([1, 2]).reshape(1, 2, 8), 2, [1])           # host polynomials: device only


def test_merkle_paths_vs_single_path(ctx):
    F, h = fields.PALLAS_FQ, hashes.keccak256
    polys = [fields.random_elements(F, 256, 3 + i) for i in range(2)]
    levels, _ = fri.precommit(polys, F, 256, 2, h)
    tree = ctx.merkle_commit(F.name, 0, dev(to_arr([v for c in polys for v in c]).reshape(2, 256, 8)), 8, 2, keep_tree=True)
    idx = [0, 1, 63, 17, 17, 40]
    got = tree.paths(idx)
    assert got == [tree.path(i) for i in idx] == [fri.merkle_proof(levels, i) for i in idx]
    with pytest.raises(Exception):
        tree.paths([64])
    assert tree.paths([]) == []


@pytest.mark.parametrize("F,hid,steps,degree_log,expand,grind,sizes", [
    (fields.PALLAS_FQ, 0, [1, 1, 1], 4, 2, False, (16, 16, 16)),
    (fields.PALLAS_FP, 2, [2, 1, 1], 5, 2, True, (32, 32, 16)),
    (fields.BLS12_381_FR, 1, [3, 1], 5, 2, True, (32, 128, 32)),
    (fields.BN254_FR, 0, [1, 2, 2, 1], 7, 1, False, (128, 64, 128)),
], ids=["steps111", "steps211-grind-k512", "steps31-grind-sha-d0poly", "steps1221"])
def test_lpc_full_proof_vs_oracle(ctx, F, hid, steps, degree_log, expand, grind, sizes):
    """Complete lpc proof (lpc.hpp:113-200 -> basic_fri.hpp:670-923: commit phase, grinding, query phase) from the
    device path: identical to the oracle prover's proof, accepted by the oracle verifier (basic_fri.hpp:932-1150,
    lpc.hpp:202-263), rejected once tampered.  Batches mix polynomial sizes, one batch may already live on D[0]."""
    from crypto3_zk_b200.lpc import FriParams, LpcCommitmentScheme
    from oracle import fri_query
    h, p, lam = HASHES[hid], F.p, 5
    d0 = 1 << (degree_log + expand)
    def rnd(size, seed):   # evaluations of a polynomial of degree <= max_degree, possibly on a larger domain
        n0 = min(size, 1 << degree_log)
        v = fields.random_elements(F, n0, seed)
        return v if n0 == size else ntt.dfs_resize(v, F, size)

    polys = {0: [rnd(sizes[0], 1 + i) for i in range(2)],
             1: [rnd(sizes[1], 10 + i) for i in range(3)],
             4: [rnd(sizes[2], 20)]}
    y = fields.random_elements(F, 1, 77)[0]
    yw = y * F.omega(degree_log) % p
    points = {0: [[y], [y]], 1: [[y], [y, yw], [yw]], 4: [[y]]}
    params = fri_query.FriParams(F, steps, degree_log, lam, expand, grind, 0x3FF)
    tr = hashes.FiatShamirSequential(h, b"\x07")
    etha = tr.challenge(F)
    trees = {k: fri.precommit(polys[k], F, d0, steps[0], h)[0] for k in polys}
    fixed_values = {0: [lpc.poly_eval(ntt.dfs_coefficients(q, F), etha, p) for q in polys[0]]}
    tp = tr.copy()
    want = fri_query.lpc_proof_eval(polys, points, trees, params, tp, h, (0,), etha, fixed_values)
    # ---- device
    # every other case keeps the extended evaluations on the device (query phase = gather): same proof either way
    scheme = LpcCommitmentScheme(ctx, F.name, hid, FriParams(steps, degree_log, lam, expand, grind, 0x3FF),
                                 retain_lde=bool(len(steps) % 2))
    t2 = OracleTranscript(h, F, b"\x07")
    for k in polys:
        nk = len(polys[k][0])
        scheme.append_to_batch(k, dev(to_arr([v for c in polys[k] for v in c]).reshape(len(polys[k]), nk, 8)))
    scheme.mark_batch_as_fixed(0)
    scheme.setup(t2, fixed_values)
    commitments = {k: scheme.commit(k) for k in polys}
    assert commitments == {k: trees[k][-1][0] for k in trees}
    for k in points:
        for i, pts in enumerate(points[k]):
            for x in pts:
                scheme.append_eval_point(k, x, poly=i)
    got = scheme.proof_eval(t2, query=True)["proof"]
    assert got["z"] == want["z"]
    gf, wf = got["fri_proof"], want["fri_proof"]
    assert gf["fri_roots"] == wf["fri_roots"] and gf["final_polynomial"] == wf["final_polynomial"]
    assert gf["proof_of_work"] == wf["proof_of_work"]
    assert gf["query_proofs"] == wf["query_proofs"]
    assert t2.state == tp.state
    tv = tr.copy()
    assert fri_query.lpc_verify_eval(got, points, commitments, params, tv, h, (0,), etha, fixed_values)
    bad = {"z": got["z"], "fri_proof": dict(gf, final_polynomial=[(gf["final_polynomial"][0] + 1) % p] + gf["final_polynomial"][1:])}
    assert not fri_query.lpc_verify_eval(bad, points, commitments, params, tr.copy(), h, (0,), etha, fixed_values)


def test_lpc_full_size_proof_verifies(ctx):
    """BASELINE configs[4] size: 2^20-row Pallas columns, blow-up 8, 19 FRI rounds of step 1 (fri_params(1, 20, lambda, 3),
    test/systems/plonk/placeholder/placeholder.cpp:231), keccak-256.  The complete lpc proof of the device path is
    accepted by the oracle verifier (basic_fri.hpp:932-1150 / lpc.hpp:202-263), which never sees the polynomials, and
    rejected after one opened value changes.  lambda and the column counts are kept small only because the verifier's
    Keccak is pure Python."""
    import torch
    from crypto3_zk_b200.lpc import FriParams, LpcCommitmentScheme
    from crypto3_zk_b200.transcript import FiatShamirSequential
    from oracle import fri_query
    F, h, hid, lam, degree_log, expand = fields.PALLAS_FP, hashes.keccak256, 0, 6, 20, 3
    n, p = 1 << degree_log, fields.PALLAS_FP.p
    counts = {0: 5, 1: 4, 2: 1, 3: 2}
    g = torch.Generator(device="cuda").manual_seed(5)
    cols = {}
    for k, cnt in counts.items():
        x = torch.randint(-2**31, 2**31 - 1, (cnt, n, 8), dtype=torch.int32, device="cuda", generator=g)
        x[..., 7] &= 0x0FFFFFFF
        cols[k] = x
    fri = FriParams.with_max_step_one(degree_log, lam, expand, use_grinding=True, grinding_parameter=0xFFFF)
    scheme = LpcCommitmentScheme(ctx, F.name, hid, fri)
    tr = FiatShamirSequential(hid, b"full-size")
    for k in counts:
        scheme.append_to_batch(k, cols[k])
    scheme.mark_batch_as_fixed(0)
    scheme.commit(0)
    scheme.setup(tr, None)
    commitments = {0: scheme._trees[0].root()}
    for k in (1, 2, 3):
        commitments[k] = scheme.commit(k)
    y = 0x1234567890abcdef1234567890abcdef % p
    yw = y * F.omega(degree_log) % p
    for k in counts:
        scheme.append_eval_point(k, y)
    for k in (1, 2):
        scheme.append_eval_point(k, yw)
    co = ctx.ntt(F.name, cols[0].clone(), degree_log, inverse=True)
    scheme._fixed_values = {0: [v[0] for v in ctx.poly_evaluate(F.name, co, n, [scheme._etha])]}
    etha, fixed_values = scheme._etha, scheme._fixed_values
    t_prover = FiatShamirSequential(hid, b"full-size")
    t_prover.state = tr.state
    res = scheme.proof_eval(t_prover, query=True)
    proof = res["proof"]
    assert all(r == 0 for r in res["remainders"])
    assert len(proof["fri_proof"]["query_proofs"]) == lam and len(proof["fri_proof"]["fri_roots"]) == degree_log - 1
    points = {k: [list(pts) for pts in scheme._points[k]] for k in counts}
    params = fri_query.FriParams(F, fri.step_list, degree_log, lam, expand, True, 0xFFFF)

    def verifier_transcript():
        t = hashes.FiatShamirSequential(h, b"full-size")
        t.state = tr.state
        return t

    tv = verifier_transcript()
    assert fri_query.lpc_verify_eval(proof, points, commitments, params, tv, h, (0,), etha, fixed_values)
    assert tv.state == t_prover.state
    # and by the product's own host verifier (lpc.hpp:202-263), through the scheme object
    t_own = FiatShamirSequential(hid, b"full-size")
    t_own.state = tr.state
    assert scheme.verify_eval(proof, commitments, t_own) and t_own.state == t_prover.state
    q0 = proof["fri_proof"]["query_proofs"][0]
    v = q0["initial_proof"][1]["values"][2][0]
    v[1] = (v[1] + 1) % p
    assert not fri_query.lpc_verify_eval(proof, points, commitments, params, verifier_transcript(), h, (0,), etha, fixed_values)


# ------------------------------------------------------------------------------------------ permutation argument (8(f)-3)
@pytest.mark.parametrize("F", [fields.PALLAS_FP, fields.BLS12_381_FR], ids=lambda f: f.name)
@pytest.mark.parametrize("n", [1, 2, 7, 8, 9, 2047, 2048, 2049, 5000, 1 << 16])
def test_prefix_product_and_batch_inverse_vs_oracle(ctx, F, n):
    from crypto3_zk_b200 import capi
    from oracle import placeholder
    x = fields.random_elements(F, n, 3 + n)
    d = dev(to_arr(x))
    assert from_arr(host(ctx.prefix_product(F.name, d))) == placeholder.prefix_product(x, F.p)
    assert from_arr(host(ctx.prefix_product(F.name, d, exclusive=False))) == placeholder.prefix_product(x, F.p, exclusive=False)
    inv = from_arr(host(ctx.batch_inverse(F.name, d)))
    if n <= 5000:
        assert inv == placeholder.batch_inverse(x, F.p)
    else:
        assert all(a * b % F.p == 1 for a, b in zip(x[::97], inv[::97]))
    if n >= 2:
        x[n // 2] = 0
        with pytest.raises(capi.ZkbInvalidArgument):
            ctx.batch_inverse(F.name, dev(to_arr(x)))
        got = from_arr(host(ctx.prefix_product(F.name, dev(to_arr(x)), exclusive=False)))      # zeros are fine for products
        assert got[n // 2:] == [0] * (n - n // 2)


@pytest.mark.parametrize("n,ncols", [(1, 1), (8, 2), (2048, 3), (3000, 1), (1 << 14, 4)])
def test_permutation_grand_product_vs_oracle(ctx, n, ncols):
    """permutation_argument.hpp:104-133 bit for bit; the last case also checks that the product closes for columns
    that satisfy the copy constraints of a permutation of their cells (the relation the argument proves)"""
    import random
    from oracle import placeholder
    F = fields.PALLAS_FP
    p = F.p
    cols = [fields.random_elements(F, n, 11 + i) for i in range(ncols)]
    sid = [fields.random_elements(F, n, 21 + i) for i in range(ncols)]
    ssg = [fields.random_elements(F, n, 31 + i) for i in range(ncols)]
    beta, gamma = fields.random_elements(F, 2, 5)
    if n == 1 << 14:
        rnd = random.Random(9)
        cells = [(i, j) for i in range(ncols) for j in range(n)]
        perm = cells[:]
        rnd.shuffle(perm)
        sigma = dict(zip(cells, perm))
        val, seen = {}, set()
        for c in cells:
            if c in seen:
                continue
            v, x = rnd.randrange(p), c
            while x not in seen:
                seen.add(x)
                val[x] = v
                x = sigma[x]
        cols = [[val[(i, j)] for j in range(n)] for i in range(ncols)]
        ssg = [[sid[sigma[(i, j)][0]][sigma[(i, j)][1]] for j in range(n)] for i in range(ncols)]

    def t(v):
        return dev(to_arr([x for c in v for x in c]).reshape(ncols, n, 8))

    got = from_arr(host(ctx.permutation_grand_product(F.name, t(cols), t(sid), t(ssg), beta, gamma)))
    assert got == placeholder.permutation_grand_product(cols, sid, ssg, beta, gamma, F)
    if n == 1 << 14:
        nom = denom = 1
        for i in range(ncols):
            nom = nom * (cols[i][n - 1] + beta * sid[i][n - 1] + gamma) % p
            denom = denom * (cols[i][n - 1] + beta * ssg[i][n - 1] + gamma) % p
        assert got[n - 1] * nom % p * pow(denom, p - 2, p) % p == 1


@pytest.mark.parametrize("n,usable,counts", [(8, 5, (1, 1, 2)), (64, 63, (2, 1, 3)), (4096, 4000, (3, 2, 5)), (1 << 14, (1 << 14) - 3, (1, 1, 2))])
def test_lookup_grand_product_vs_oracle(ctx, n, usable, counts):
    """compute_V_L (lookup_argument.hpp:375-409) bit for bit, including the zero tail beyond the usable rows"""
    from oracle import placeholder
    F = fields.BLS12_381_FR
    ni, nv, ns = counts
    inp = [fields.random_elements(F, n, 41 + i) for i in range(ni)]
    val = [fields.random_elements(F, n, 51 + i) for i in range(nv)]
    srt = [fields.random_elements(F, n, 61 + i) for i in range(ns)]
    beta, gamma = fields.random_elements(F, 2, 8)

    def t(v):
        return dev(to_arr([x for c in v for x in c]).reshape(len(v), n, 8))

    got = from_arr(host(ctx.lookup_grand_product(F.name, t(inp), t(val), t(srt), beta, gamma, usable)))
    want = placeholder.lookup_grand_product(inp, val, srt, beta, gamma, usable, F)
    assert got == want and got[usable + 1:] == [0] * (n - usable - 1)


@pytest.mark.parametrize("cname", ["bls12_381_g1", "bn254_g1", "pallas", "bn254_g2", "bls12_381_g2"])
def test_batch_exp_vs_oracle(ctx, cname):
    """algebra::batch_exp / windowed_exp of the Groth16 generator (generator.hpp:167-225): scalar * base for many scalars,
    affine, bit for bit; zero, one, p-1 and single-byte scalars included; host and device buffers"""
    from oracle import curves
    C = {"bls12_381_g1": curves.BLS12_381_G1, "bn254_g1": curves.BN254_G1, "pallas": curves.PALLAS,
         "bn254_g2": curves.BN254_G2, "bls12_381_g2": curves.BLS12_381_G2}[cname]
    r = C.scalar_field.p
    sc = [0, 1, 2, r - 1, 255, 256, 1 << 248, (1 << 64) - 1] + fields.random_elements(C.scalar_field, 40, 3)
    base = C.mul(C.gen, 5)
    want = [C.mul(base, k) for k in sc]
    cl = C.coord_limbs32
    for arr in (to_arr(sc), dev(to_arr(sc))):
        got = ctx.batch_exp(cname, base, arr)
        g = host(got) if not isinstance(got, np.ndarray) else got
        from crypto3_zk_b200.api import _affine_from_limbs
        deg = 2 if cname.endswith("g2") else 1
        pts = [_affine_from_limbs(g[i].reshape(-1), cl, deg) for i in range(len(sc))]
        assert pts == want


# ------------------------------------------------------------------------------------------ Groth16 (config #4)
@pytest.mark.parametrize("F", [fields.BN254_FR, fields.BLS12_381_FR], ids=lambda f: f.name)
def test_sparse_matvec_vs_oracle(ctx, F):
    """cs.constraints[i].a/b/c.evaluate(full_variable_assignment) (r1cs_to_qap.hpp:245-248, 289-291), including the
    closing constraint of generate_r1cs_example_with_field_input whose rows sum every variable (long-row path)."""
    import torch
    from crypto3_zk_b200.groth16 import R1csConstraintSystem
    from oracle import groth16
    p = F.p
    cs, primary, aux = groth16.example_with_field_input(F, 5000, 10, seed=3)
    x = [1] + primary + aux
    pcs = R1csConstraintSystem(cs.num_inputs, cs.num_aux, cs.constraints)
    xd = dev(to_arr(x))
    for side in range(3):
        rp, ci, va = pcs.csr(side)
        mat = ctx.sparse_matrix(F.name, cs.num_constraints, cs.num_variables + 1, rp, ci, va)
        y = torch.zeros((cs.num_constraints, 8), dtype=torch.int32, device="cuda")
        mat.matvec(xd, y)
        want = [sum(co * x[i] for i, co in con[side]) % p for con in cs.constraints]
        assert from_arr(host(y)) == want
        y2 = torch.zeros_like(y)
        mat.matvec(to_arr(x), y2)      # host assignment
        assert torch.equal(y, y2)
        mat.free()


def _device_key(ctx, pk, G1, G2):
    from crypto3_zk_b200.groth16 import ProvingKey, R1csConstraintSystem
    pcs = R1csConstraintSystem(pk.cs.num_inputs, pk.cs.num_aux, pk.cs.constraints)
    return ProvingKey(ctx, G1.name, G2.name, pcs, pk.alpha_g1, pk.beta_g1, pk.beta_g2, pk.delta_g1, pk.delta_g2,
                      pk.A_query, pk.B_indices, pk.B_g2, pk.B_g1, pk.H_query, pk.L_query)


@pytest.mark.parametrize("curve,kind,nc,ni", [("bn254", "field", 28, 3), ("bn254", "binary", 59, 4), ("bls12_381", "field", 13, 2)])
def test_groth16_prove_vs_oracle(ctx, curve, kind, nc, ni):
    """r1cs_gg_ppzksnark_prover::process (prover.hpp:73-158): witness map + the four MSMs (A, B on G2 and G1, H, L) and
    the proof assembly on the device, bit-exact (affine) against the oracle's restatement and against the Groth16
    relations in the exponent."""
    from crypto3_zk_b200 import groth16 as dg
    from oracle import curves, groth16
    F, G1, G2 = ((fields.BN254_FR, curves.BN254_G1, curves.BN254_G2) if curve == "bn254" else
                 (fields.BLS12_381_FR, curves.BLS12_381_G1, curves.BLS12_381_G2))
    make = groth16.example_with_field_input if kind == "field" else groth16.example_with_binary_input
    cs, primary, aux = make(F, nc, ni, seed=5)
    t, alpha, beta, gamma, delta = fields.random_elements(F, 5, 21)
    pk = groth16.generator(cs, G1, G2, F, t, alpha, beta, gamma, delta)
    r, s = fields.random_elements(F, 2, 22)
    want = groth16.prove(pk, primary, aux, r, s, G1, G2, F)
    dpk = _device_key(ctx, pk, G1, G2)
    # the witness map alone (r1cs_to_qap.hpp:219-325)
    m, full, H = groth16.witness_map(pk.cs, primary, aux, F)
    h = dg.witness_map(ctx, dpk, dev(to_arr([1] + full)))
    assert from_arr(host(h)) == H[:m]
    got = dg.prove(ctx, dpk, primary, aux, r, s)
    assert got == want
    assert dg.prove(ctx, dpk, primary, aux, r, s, concurrent=True) == want     # five multiexps on five streams
    a, b, c = groth16.proof_in_the_exponent(pk, primary, aux, r, s, F)
    assert got[0] == G1.mul(G1.gen, a) and got[1] == G2.mul(G2.gen, b) and got[2] == G1.mul(G1.gen, c)
    # zero randomness and a binary assignment exercise the 0/1 scalar paths of the MSM
    assert dg.prove(ctx, dpk, primary, aux, 0, 0) == groth16.prove(pk, primary, aux, 0, 0, G1, G2, F)
    if curve == "bls12_381":
        # the reference's wire format on both sides of the prover (r1cs_gg_ppzksnark/marshalling.hpp): proving key read
        # from its byte blob, proof written as the 192-byte g_A | g_B | g_C blob
        from crypto3_zk_b200 import marshalling
        from crypto3_zk_b200.groth16 import ProvingKey, R1csConstraintSystem
        blob = marshalling.proving_key_to_bytes(dict(
            alpha_g1=pk.alpha_g1, beta_g1=pk.beta_g1, beta_g2=pk.beta_g2, delta_g1=pk.delta_g1, delta_g2=pk.delta_g2,
            A_query=pk.A_query, B_indices=pk.B_indices, B_g2=pk.B_g2, B_g1=pk.B_g1, B_domain_size=len(pk.A_query),
            H_query=pk.H_query, L_query=pk.L_query, num_inputs=pk.cs.num_inputs, num_aux=pk.cs.num_aux,
            constraints=pk.cs.constraints))
        k = marshalling.proving_key_from_bytes(blob)
        dpk2 = ProvingKey(ctx, G1.name, G2.name, R1csConstraintSystem(k["num_inputs"], k["num_aux"], k["constraints"]),
                          k["alpha_g1"], k["beta_g1"], k["beta_g2"], k["delta_g1"], k["delta_g2"], k["A_query"],
                          k["B_indices"], k["B_g2"], k["B_g1"], k["H_query"], k["L_query"])
        wire = marshalling.proof_to_bytes(dg.prove(ctx, dpk2, primary, aux, r, s))
        assert len(wire) == 192 and wire == marshalling.proof_to_bytes(want)
        assert marshalling.proof_from_bytes(wire) == want
        # the same blob with the query vectors decompressed on the device (zkb_points_decompress): same points, same proof
        kd = marshalling.proving_key_from_bytes(blob, ctx=ctx)
        from crypto3_zk_b200.api import _affine_from_limbs
        for name, cl, deg in (("A_query", 12, 1), ("B_g1", 12, 1), ("H_query", 12, 1), ("L_query", 12, 1), ("B_g2", 24, 2)):
            arr = kd[name].cpu().numpy().view(np.uint32)
            assert [_affine_from_limbs(arr[i].reshape(-1), cl, deg) for i in range(arr.shape[0])] == list(k[name]), name
        assert list(kd["B_indices"]) == list(k["B_indices"]) and kd["B_domain_size"] == k["B_domain_size"]
        dpk3 = ProvingKey(ctx, G1.name, G2.name, R1csConstraintSystem(kd["num_inputs"], kd["num_aux"], kd["constraints"]),
                          kd["alpha_g1"], kd["beta_g1"], kd["beta_g2"], kd["delta_g1"], kd["delta_g2"], kd["A_query"],
                          kd["B_indices"], kd["B_g2"], kd["B_g1"], kd["H_query"], kd["L_query"])
        assert marshalling.proof_to_bytes(dg.prove(ctx, dpk3, primary, aux, r, s)) == wire


@pytest.mark.parametrize("curve,nc,ni", [("bls12_381", 13, 2), ("bn254", 27, 4)])
def test_groth16_generator_vs_oracle(ctx, curve, nc, ni):
    """r1cs_gg_ppzksnark_generator::basic_process (generator.hpp:83-235): QAP evaluation on the host, every key element by
    zkb_batch_exp; same key as the oracle's generator, and the device prover proves with it."""
    from crypto3_zk_b200 import groth16 as dg
    from oracle import curves, groth16
    F, G1, G2 = ((fields.BN254_FR, curves.BN254_G1, curves.BN254_G2) if curve == "bn254" else
                 (fields.BLS12_381_FR, curves.BLS12_381_G1, curves.BLS12_381_G2))
    cs, primary, aux = groth16.example_with_field_input(F, nc, ni, seed=7)
    t, alpha, beta, gamma, delta = fields.random_elements(F, 5, 31)
    pk = groth16.generator(cs, G1, G2, F, t, alpha, beta, gamma, delta)
    pcs = dg.R1csConstraintSystem(cs.num_inputs, cs.num_aux, list(cs.constraints))
    key, vk = dg.generator(ctx, G1.name, G2.name, pcs, t, alpha, beta, gamma, delta)
    for name in ("alpha_g1", "beta_g1", "beta_g2", "delta_g1", "delta_g2", "A_query", "B_indices", "B_g2", "B_g1",
                 "H_query", "L_query"):
        assert key[name] == getattr(pk, name), name
    assert vk["gamma_g2"] == G2.mul(G2.gen, gamma) and vk["delta_g2"] == pk.delta_g2
    r, s = fields.random_elements(F, 2, 32)
    dpk = dg.proving_key_from_dict(ctx, G1.name, G2.name, key)
    assert dg.prove(ctx, dpk, primary, aux, r, s) == groth16.prove(pk, primary, aux, r, s, G1, G2, F)
    # the same generator with the QAP evaluation and every key vector on the device (generator_device)
    from crypto3_zk_b200.api import _affine_from_limbs
    cs2 = dg.swap_ab_if_beneficial(pcs)
    dkey, dvk = dg.generator_device(ctx, G1.name, G2.name, cs2.num_constraints, cs2.num_inputs, cs2.num_variables,
                                    [cs2.csr(k) for k in range(3)], t, alpha, beta, gamma, delta)

    def pts(tensor, curve):
        a = tensor.cpu().numpy().view(np.uint32)
        return [_affine_from_limbs(a[i].reshape(-1), curve.coord_limbs32, 2 if curve is G2 else 1) for i in range(a.shape[0])]

    for name in ("alpha_g1", "beta_g1", "beta_g2", "delta_g1", "delta_g2"):
        assert dkey[name] == getattr(pk, name), name
    for name, curve in (("A_query", G1), ("B_g1", G1), ("H_query", G1), ("L_query", G1), ("B_g2", G2)):
        assert pts(dkey[name], curve) == list(getattr(pk, name)), name
    assert list(dkey["B_indices"]) == list(pk.B_indices)
    assert pts(dvk["gamma_ABC_g1"], G1) == [vk["gamma_ABC_g1"][0]] + list(vk["gamma_ABC_g1"][1]) and dvk["gamma_g2"] == vk["gamma_g2"]
    dkey["constraints"] = cs2.constraints
    dpk2 = dg.proving_key_from_dict(ctx, G1.name, G2.name, dkey)
    assert dg.prove(ctx, dpk2, primary, aux, r, s) == groth16.prove(pk, primary, aux, r, s, G1, G2, F)


def test_qap_instance_evaluation_device_at_2p12(ctx):
    """instance_map_with_evaluation (r1cs_to_qap.hpp:138-204) on the device against the host evaluation on Python integers,
    4093 constraints from the numpy CSR builder of the 2^22 workload (transposed mat-vec, batched inversion, product scans)"""
    from crypto3_zk_b200 import groth16 as dg
    from crypto3_zk_b200 import workloads as W
    F = fields.BN254_FR
    ni = 2
    nc = (1 << 12) - ni - 1
    shape, sides, full = W.groth16_field_input_example(F.name, nc, ni, seed=3)
    t = fields.random_elements(F, 1, 77)[0]
    at, bt, ct, ht, zt, m = dg.qap_instance_evaluation_device(ctx, dg.FIELD_BY_NAME[F.name], nc, ni, shape.num_variables, sides, t)
    cons = []
    for row in range(nc):
        con = []
        for row_ptr, col, vals in sides:
            lo, hi = int(row_ptr[row]), int(row_ptr[row + 1])
            con.append([(int(col[k]), fields.u32_array_to_ints(vals[k:k + 1])[0]) for k in range(lo, hi)])
        cons.append(tuple(con))
    host_cs = dg.R1csConstraintSystem(ni, shape.num_aux, cons)
    At, Bt, Ct, Ht, Zt, mm = dg.qap_instance_evaluation(host_cs, dg.FIELD_BY_NAME[F.name], t)
    assert (zt, m) == (Zt, mm)
    assert from_arr(host(at)) == At and from_arr(host(bt)) == Bt and from_arr(host(ct)) == Ct and from_arr(host(ht)) == Ht
