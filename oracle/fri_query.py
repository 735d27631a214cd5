"""FRI query phase, grinding and verifier (oracle; test infrastructure only - nothing under crypto3_zk_b200/ imports it).

Literal restatement, in plain Python integers, of
  get_paired_index / get_folded_index / make_proof_specialized   basic_fri.hpp:349-351, 526-542
  calculate_s, get_correct_order                                 basic_fri.hpp:583-668
  proof_eval, grinding + "Query phase"                           basic_fri.hpp:743-923
  verify_eval (FRI)                                              basic_fri.hpp:932-1150
  proof_of_work<Hash, uint32>::generate / verify                 proof_of_work.hpp:47-81
  lpc_commitment_scheme::proof_eval / verify_eval                zk/commitments/polynomial/lpc.hpp:113-200, 202-263
all under zk/commitments/detail/polynomial/ unless a full path is given.  Parity status: "unpinned" by reference
fixtures (the reference holds no concrete LPC proof); the pin is the scheme relation itself - a proof produced by the
restated prover (and by the CUDA path) must pass the restated verifier, and a tampered proof must not.

Data model
  merkle proof : {"index": leaf index, "path": [sibling digests, leaf level first], "root": bytes}
  fri proof    : {"fri_roots": [..], "final_polynomial": [coefficients], "proof_of_work": int or None,
                  "query_proofs": [{"initial_proof": {k: {"values": [poly][j][2], "p": merkle proof}},
                                    "round_proofs": [{"y": [j][2], "p": merkle proof}]}]}
  lpc proof    : {"z": {k: [poly][point]}, "fri_proof": fri proof}
"""
from . import fri as _fri
from .lpc import combined_q, eval_polys, poly_eval, unique_points
from .ntt import EvaluationDomain, dfs_coefficients


class FriParams:
    """basic_batched_fri::params_type (basic_fri.hpp:151-229): D[i] has size 2^(degree_log + expand_factor - i)."""

    def __init__(self, field, step_list, degree_log, lambda_, expand_factor, use_grinding=False, grinding_parameter=0xFFFF):
        self.field = field
        self.step_list = list(step_list)
        self.r = sum(step_list)
        self.lambda_ = lambda_
        self.expand_factor = expand_factor
        self.max_degree = (1 << degree_log) - 1
        self.use_grinding = use_grinding
        self.grinding_parameter = grinding_parameter
        self.log_d0 = degree_log + expand_factor
        self.D = [EvaluationDomain(field, 1 << (self.log_d0 - i)) for i in range(self.r + 1)]


def check_step_list(params):
    """basic_fri.hpp:544-571"""
    if not params.step_list:
        return False
    for s in params.step_list:
        if s <= 0 or s > 10:
            return False
    return sum(params.step_list) == params.r and params.step_list[-1] == 1


def get_paired_index(x_index, domain_size):
    return (x_index + domain_size // 2) % domain_size


def get_folded_index(x_index, domain_size, fri_step):
    for _ in range(fri_step):
        domain_size //= 2
        x_index %= domain_size
    return x_index


def calculate_s(x, x_index, fri_step, D):
    """basic_fri.hpp:583-617 (m = 2)"""
    domain_size = D.m
    coset_size = 1 << fri_step
    s = [[0, 0] for _ in range(coset_size // 2)]
    s_indices = [[0, 0] for _ in range(coset_size // 2)]
    s_indices[0][0] = x_index
    s_indices[0][1] = get_paired_index(x_index, domain_size)
    s[0][0] = D.get_domain_element(s_indices[0][0])
    s[0][1] = D.get_domain_element(s_indices[0][1])
    assert s[0][0] == x
    base_index = domain_size // 4
    prev_half_size = 1
    i = 1
    while i < coset_size // 2:
        for j in range(prev_half_size):
            s_indices[i][0] = (base_index + s_indices[j][0]) % domain_size
            s_indices[i][1] = get_paired_index(s_indices[i][0], domain_size)
            s[i][0] = D.get_domain_element(s_indices[i][0])
            s[i][1] = D.get_domain_element(s_indices[i][1])
            i += 1
        base_index //= 2
        prev_half_size <<= 1
    return s, s_indices


def get_correct_order(x_index, domain_size, fri_step, input_s_indices):
    """basic_fri.hpp:619-668: [(position in input_s_indices, 0 straight / 1 swapped)] in leaf order"""
    coset_size = 1 << fri_step
    assert coset_size // 2 == len(input_s_indices)
    ordered = [0] * (coset_size // 2)
    ordered[0] = get_folded_index(x_index, domain_size, fri_step)
    base_index = domain_size // 4
    prev_half_size = 1
    i = 1
    while i < coset_size // 2:
        for j in range(prev_half_size):
            ordered[i] = (base_index + ordered[j]) % domain_size
            i += 1
        base_index //= 2
        prev_half_size <<= 1
    out = []
    for i in range(coset_size // 2):
        paired = get_paired_index(ordered[i], domain_size)
        found = None
        for pos, v in enumerate(input_s_indices):
            if v[0] == ordered[i] and v[1] == paired:
                found = (pos, 0)
                break
            if v[1] == ordered[i] and v[0] == paired:
                found = (pos, 1)
                break
        assert found is not None
        out.append(found)
    return out


# ------------------------------------------------------------------------------------------------ merkle proofs
def make_merkle_proof(levels, leaf_idx):
    return {"index": leaf_idx, "path": _fri.merkle_proof(levels, leaf_idx), "root": levels[-1][0]}


def make_proof_specialized(x_index, domain_size, levels):
    """basic_fri.hpp:526-531"""
    return make_merkle_proof(levels, min(x_index, get_paired_index(x_index, domain_size)))


def merkle_validate(proof, leaf_data, h):
    """containers::merkle_proof<Hash,2>::validate (upstream; believed: leaf digest = H(data), node = H(left || right))"""
    d = h(leaf_data)
    idx = proof["index"]
    for sib in proof["path"]:
        d = h(sib + d) if idx & 1 else h(d + sib)
        idx >>= 1
    return d == proof["root"]


# ------------------------------------------------------------------------------------------------ grinding
def _be32(v):
    return bytes([(v >> 24) & 0xFF, (v >> 16) & 0xFF, (v >> 8) & 0xFF, v & 0xFF])


def pow_generate(transcript, mask=0xFFFF, start=0):
    """proof_of_work.hpp:47-68; `start` replaces std::rand() (any passing nonce verifies)"""
    pow_ = start & 0xFFFFFFFF
    while True:
        tmp = transcript.copy()
        tmp.absorb(_be32(pow_))
        if tmp.int_challenge(32) & mask == 0:
            break
        pow_ = (pow_ + 1) & 0xFFFFFFFF
    transcript.absorb(_be32(pow_))
    transcript.int_challenge(32)
    return pow_


def pow_verify(transcript, pow_, mask=0xFFFF):
    """proof_of_work.hpp:70-78"""
    transcript.absorb(_be32(pow_))
    return transcript.int_challenge(32) & mask == 0


# ------------------------------------------------------------------------------------------------ prover
def _domain_index(D, x):
    """index of x in D.  The reference searches linearly (basic_fri.hpp:780-786); same result bit by bit for the
    2^k-th roots of unity: bit i of the exponent is set iff (x w^-e)^(2^(k-1-i)) != 1 for the bits e found so far."""
    p, k = D.field.p, D.log_m
    w_inv = D.omega_inv
    e = 0
    for i in range(k):
        if pow(x * pow(w_inv, e, p) % p, 1 << (k - 1 - i), p) != 1:
            e |= 1 << i
    if pow(D.omega, e, p) != x % p:
        raise ValueError("x is not in the domain")
    return e


def fri_proof_eval(g, combined_Q, precommitments, params, transcript, h):
    """zk::algorithms::proof_eval<FRI> (basic_fri.hpp:670-923).
    g: {k: [dfs value lists]}; combined_Q: dfs on D[0]; precommitments: {k: merkle levels}."""
    F = params.field
    p = F.p
    assert check_step_list(params)
    # ---- commit phase (:706-742)
    cp = _fri.commit_phase(combined_Q, F, params.log_d0, params.step_list, h, transcript)
    fs, fri_levels, final_polynomial = cp["fs"], cp["levels"], cp["final_polynomial"]
    proof = {"fri_roots": cp["roots"], "final_polynomial": final_polynomial, "proof_of_work": None}
    # ---- grinding (:744-747)
    if params.use_grinding:
        proof["proof_of_work"] = pow_generate(transcript, params.grinding_parameter)
    # ---- query phase (:749-915)
    d0 = params.D[0].m
    g_coeffs = {k: [None if len(poly) == d0 else dfs_coefficients(poly, F) for poly in polys] for k, polys in g.items()}
    query_proofs = []
    for _ in range(params.lambda_):
        domain_size = d0
        x = transcript.challenge(F)
        x = pow(x, (p - 1) // domain_size, p)
        x_index = _domain_index(params.D[0], x)
        s, s_indices = calculate_s(x, x_index, params.step_list[0], params.D[0])
        initial_proof = {}
        coset_size = 1 << params.step_list[0]
        for k in sorted(g):
            values = []
            for pi, poly in enumerate(g[k]):
                vals = []
                for j in range(coset_size // 2):
                    if len(poly) == d0:
                        ind0, ind1 = min(s_indices[j]), max(s_indices[j])
                        vals.append([poly[ind0], poly[ind1]])
                    else:
                        s0, s1 = (s[j][0], s[j][1]) if s_indices[j][0] < s_indices[j][1] else (s[j][1], s[j][0])
                        vals.append([poly_eval(g_coeffs[k][pi], s0, p), poly_eval(g_coeffs[k][pi], s1, p)])
                values.append(vals)
            initial_proof[k] = {"values": values,
                                "p": make_proof_specialized(get_folded_index(x_index, d0, params.step_list[0]), d0,
                                                            precommitments[k])}
        round_proofs = []
        t = 0
        for i in range(len(params.step_list)):
            domain_size = params.D[t].m
            x_index %= domain_size
            x = params.D[t].get_domain_element(x_index)
            rp = {"p": make_proof_specialized(get_folded_index(x_index, domain_size, params.step_list[i]), domain_size,
                                              fri_levels[i])}
            t += params.step_list[i]
            if i < len(params.step_list) - 1:
                x_index %= params.D[t].m
                x = params.D[t].get_domain_element(x_index)
                s, s_indices = calculate_s(x, x_index, params.step_list[i + 1], params.D[t])
                cs = 1 << params.step_list[i + 1]
                rp["y"] = [[fs[i + 1][min(s_indices[j])], fs[i + 1][max(s_indices[j])]] for j in range(cs // 2)]
            else:
                x_index %= params.D[t - 1].m
                x = params.D[t - 1].get_domain_element(x_index)
                x = x * x % p
                ind = 0 if x_index % (params.D[t - 1].m // 2) < params.D[t - 1].m // 4 else 1
                y = [[0, 0]]
                y[0][ind] = poly_eval(final_polynomial, x, p)
                y[0][1 - ind] = poly_eval(final_polynomial, (p - x) % p, p)
                rp["y"] = y
            round_proofs.append(rp)
        query_proofs.append({"initial_proof": initial_proof, "round_proofs": round_proofs})
    proof["query_proofs"] = query_proofs
    return proof


# ------------------------------------------------------------------------------------------------ verifier
def _leaf_data(elems, field):
    return b"".join(int(v).to_bytes(field.nbytes, "big") for v in elems)


def _interp2(s_ch, y0, y1, alpha, p):
    """lagrange_interpolation{(s, y0), (-s, y1)} evaluated at alpha"""
    inv = pow((2 * s_ch) % p, p - 2, p)
    return (y0 + (y0 - y1) * (alpha - s_ch) % p * inv) % p


def fri_verify_eval(proof, params, commitments, theta, poly_ids, combined_U, denominators, transcript, h):
    """zk::algorithms::verify_eval<FRI> (basic_fri.hpp:932-1150); denominators: points (V = X - point)."""
    F = params.field
    p = F.p
    assert check_step_list(params)
    assert len(combined_U) == len(denominators) == len(poly_ids)
    fp = proof["final_polynomial"]
    deg = max((i for i, v in enumerate(fp) if v), default=0)
    log_md = (params.max_degree + 1).bit_length() - 1
    if deg > 2 ** (log_md - params.r + 1) - 1:
        return False
    alphas = []
    for i, step in enumerate(params.step_list):
        transcript.absorb(proof["fri_roots"][i])
        for _ in range(step):
            alphas.append(transcript.challenge(F))
    if params.use_grinding and not pow_verify(transcript, proof["proof_of_work"], params.grinding_parameter):
        return False
    for query_id in range(params.lambda_):
        qp = proof["query_proofs"][query_id]
        domain_size = params.D[0].m
        coset_size = 1 << params.step_list[0]
        x = pow(transcript.challenge(F), (p - 1) // domain_size, p)
        x_index = _domain_index(params.D[0], x)
        s, s_indices = calculate_s(x, x_index, params.step_list[0], params.D[0])
        correct_order_idx = get_correct_order(x_index, domain_size, params.step_list[0], s_indices)
        # ---- initial proofs (:984-1004)
        for k, ip in qp["initial_proof"].items():
            if ip["p"]["root"] != commitments[k]:
                return False
            data = []
            for vals in ip["values"]:
                for idx, _pair in correct_order_idx:
                    data += [vals[idx][0], vals[idx][1]]
            if not merkle_validate(ip["p"], _leaf_data(data, F), h):
                return False
        # ---- combined Q values (:1006-1036)
        theta_acc = 1
        y = [[0, 0] for _ in range(coset_size // 2)]
        for pidx in range(len(poly_ids)):
            Q = [[0, 0] for _ in range(coset_size // 2)]
            for (bk, bi) in poly_ids[pidx]:
                for j in range(coset_size // 2):
                    Q[j][0] = (Q[j][0] + qp["initial_proof"][bk]["values"][bi][j][0] * theta_acc) % p
                    Q[j][1] = (Q[j][1] + qp["initial_proof"][bk]["values"][bi][j][1] * theta_acc) % p
                theta_acc = theta_acc * theta % p
            for j in range(coset_size // 2):
                id0 = 0 if s_indices[j][0] < s_indices[j][1] else 1
                id1 = 1 - id0
                Q[j][0] = (Q[j][0] - combined_U[pidx]) % p
                Q[j][1] = (Q[j][1] - combined_U[pidx]) % p
                Q[j][0] = Q[j][0] * pow((s[j][id0] - denominators[pidx]) % p, p - 2, p) % p
                Q[j][1] = Q[j][1] * pow((s[j][id1] - denominators[pidx]) % p, p - 2, p) % p
                y[j][0] = (y[j][0] + Q[j][0]) % p
                y[j][1] = (y[j][1] + Q[j][1]) % p
        # ---- round proofs (:1037-1135)
        t = 0
        for i, step in enumerate(params.step_list):
            coset_size = 1 << step
            rp = qp["round_proofs"][i]
            if rp["p"]["root"] != proof["fri_roots"][i]:
                return False
            s, s_indices = calculate_s(x, x_index, step, params.D[t])
            correct_order_idx = get_correct_order(x_index, domain_size, step, s_indices)
            data = []
            for idx, _pair in correct_order_idx:
                data += [y[idx][0], y[idx][1]]
            if not merkle_validate(rp["p"], _leaf_data(data, F), h):
                return False
            # colinear checks inside a multi-step round
            for _step_i in range(step - 1):
                y_next = [[0, 0] for _ in range(len(y) // 2)]
                domain_size = params.D[t].m
                x_index %= domain_size
                x = params.D[t].get_domain_element(x_index)
                _s_next, s_indices_next = calculate_s(x * x % p, x_index % params.D[t + 1].m, step, params.D[t + 1])
                s, s_indices = calculate_s(x, x_index, step, params.D[t])
                for y_ind in range(len(y_next)):
                    ind0 = 0 if s_indices[2 * y_ind][0] < s_indices[2 * y_ind][1] else 1
                    interpolant_l = _interp2(s[2 * y_ind][ind0], y[2 * y_ind][0], y[2 * y_ind][1], alphas[t], p)
                    ind0 = 0 if s_indices[2 * y_ind + 1][0] < s_indices[2 * y_ind + 1][1] else 1
                    interpolant_r = _interp2(s[2 * y_ind + 1][ind0], y[2 * y_ind + 1][0], y[2 * y_ind + 1][1], alphas[t], p)
                    if s_indices_next[y_ind][0] < s_indices_next[y_ind][1]:
                        y_next[y_ind] = [interpolant_l, interpolant_r]
                    else:
                        y_next[y_ind] = [interpolant_r, interpolant_l]
                x = x * x % p
                y = y_next
                t += 1
            domain_size = params.D[t].m
            x_index %= domain_size
            x = params.D[t].get_domain_element(x_index)
            s, s_indices = calculate_s(x, x_index, step, params.D[t])
            ind0 = 0 if s_indices[0][0] < s_indices[0][1] else 1
            interpolant = _interp2(s[0][ind0], y[0][0], y[0][1], alphas[t], p)
            ind = 0 if s_indices[0][ind0] % (params.D[t].m // 2) < params.D[t].m // 4 else 1
            if interpolant != rp["y"][0][ind]:
                return False
            y = [list(v) for v in rp["y"]]
            if i < len(params.step_list) - 1:
                t += 1
                domain_size = params.D[t].m
                x_index %= domain_size
                x = params.D[t].get_domain_element(x_index)
        # ---- final polynomial (:1137-1147)
        x_index %= params.D[t].m
        x = params.D[t].get_domain_element(x_index)
        x = x * x % p
        ind = 0 if x_index % (params.D[t].m // 2) < params.D[t].m // 4 else 1
        if y[0][ind] != poly_eval(fp, x, p):
            return False
        if y[0][1 - ind] != poly_eval(fp, (p - x) % p, p):
            return False
    return True


# ------------------------------------------------------------------------------------------------ LPC scheme
def lpc_proof_eval(polys, points, trees, params, transcript, h, fixed_batches=(), etha=None, fixed_values=None):
    """lpc_commitment_scheme::proof_eval (lpc.hpp:113-200).  polys: {k: [dfs lists]}; points: {k: [[..] per poly]};
    trees: {k: merkle levels of the committed batches}."""
    F = params.field
    z = eval_polys(polys, points, F)
    for k in sorted(trees):
        transcript.absorb(trees[k][-1][0])
    theta = transcript.challenge(F)
    _coeffs, q_dfs = combined_q(polys, points, z, theta, F, fixed_batches, etha, fixed_values)
    d0 = params.D[0].m
    q_d0 = q_dfs if len(q_dfs) == d0 else _fri.dfs_resize(q_dfs, F, d0)
    fri_proof = fri_proof_eval(polys, q_d0, trees, params, transcript, h)
    return {"z": z, "fri_proof": fri_proof}


def lpc_verify_eval(proof, points, commitments, params, transcript, h, fixed_batches=(), etha=None, fixed_values=None):
    """lpc_commitment_scheme::verify_eval (lpc.hpp:202-263)"""
    F = params.field
    p = F.p
    z = proof["z"]
    for k in sorted(commitments):
        transcript.absorb(commitments[k])
    pts = unique_points(points)
    total = len(pts) + (1 if fixed_batches else 0)
    U = [0] * total
    V = [0] * total
    poly_map = [[] for _ in range(total)]
    theta = transcript.challenge(F)
    theta_acc = 1
    for pi, point in enumerate(pts):
        V[pi] = point
        for k in sorted(z):
            for j in range(len(z[k])):
                if point not in points[k][j]:
                    continue
                U[pi] = (U[pi] + z[k][j][points[k][j].index(point)] * theta_acc) % p
                poly_map[pi].append((k, j))
                theta_acc = theta_acc * theta % p
    if total > len(pts):
        pi = len(pts)
        V[pi] = etha
        for k in sorted(z):
            if k not in fixed_batches:
                continue
            for j in range(len(z[k])):
                U[pi] = (U[pi] + fixed_values[k][j] * theta_acc) % p
                poly_map[pi].append((k, j))
                theta_acc = theta_acc * theta % p
    return fri_verify_eval(proof["fri_proof"], params, commitments, theta, poly_map, U, V, transcript, h)
