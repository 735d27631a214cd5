"""Verifier of the LPC / FRI commitment scheme on the host (CPU): lpc_commitment_scheme::verify_eval
(zk/commitments/polynomial/lpc.hpp:202-263) over zk::algorithms::verify_eval<FRI>
(zk/commitments/detail/polynomial/basic_fri.hpp:932-1150) and proof_of_work::verify (proof_of_work.hpp:70-78).

The reference's verifier is host code too: it touches lambda Merkle paths and a few field elements per query, nothing
data-parallel.  Proof objects are the ones LpcCommitmentScheme.proof_eval(query=True) returns:
  {"z": {k: [poly][point]}, "fri_proof": {"fri_roots", "final_polynomial", "proof_of_work", "query_proofs":
   [{"initial_proof": {k: {"values": [poly][j][2], "p": merkle proof}}, "round_proofs": [{"y": [j][2], "p": merkle proof}]}]}}
with merkle proof = {"index", "path" (sibling digests, leaf level first), "root"}.
"""
from .fields import FIELD_BY_NAME, omega
from .transcript import HASH_BY_ID


def _paired(x_index, domain_size):
    return (x_index + domain_size // 2) % domain_size


def _folded(x_index, domain_size, fri_step):
    for _ in range(fri_step):
        domain_size //= 2
        x_index %= domain_size
    return x_index


def _calculate_s(x_index, fri_step, log_size, F):
    """calculate_s (basic_fri.hpp:583-617): points and indices of the coset of x in D (|D| = 2^log_size)"""
    p, size = F.p, 1 << log_size
    w = omega(F, log_size)
    idx = [[x_index, _paired(x_index, size)]]
    base, prev = size // 4, 1
    while len(idx) < (1 << fri_step) // 2:
        for j in range(prev):
            a = (base + idx[j][0]) % size
            idx.append([a, _paired(a, size)])
        base //= 2
        prev <<= 1
    return [[pow(w, a, p), pow(w, b, p)] for a, b in idx], idx


def _correct_order(x_index, domain_size, fri_step, s_indices):
    """get_correct_order (basic_fri.hpp:619-668)"""
    ordered = [_folded(x_index, domain_size, fri_step)]
    base, prev = domain_size // 4, 1
    while len(ordered) < (1 << fri_step) // 2:
        for j in range(prev):
            ordered.append((base + ordered[j]) % domain_size)
        base //= 2
        prev <<= 1
    out = []
    for o in ordered:
        pr = _paired(o, domain_size)
        for pos, v in enumerate(s_indices):
            if (v[0], v[1]) == (o, pr) or (v[1], v[0]) == (o, pr):
                out.append(pos)
                break
        else:
            raise ValueError("coset index not found")
    return out


def _merkle_validate(proof, data, h, index, depth):
    """merkle_proof::validate plus what the reference leaves unchecked: the opened leaf must be the one the query selects
    (`index`) and the path must span the whole tree (`depth` levels) - without the two a prover may open another leaf,
    or pass a 64-byte inner node off as a two-element leaf (there is no leaf/node domain separation)."""
    if proof["index"] != index or len(proof["path"]) != depth:
        return False
    d, idx = h(data), proof["index"]
    for sib in proof["path"]:
        d = h(sib + d) if idx & 1 else h(d + sib)
        idx >>= 1
    return d == proof["root"]


def _leaf(elems, nbytes):
    return b"".join(int(v).to_bytes(nbytes, "big") for v in elems)


def _horner(c, x, p):
    acc = 0
    for v in reversed(c):
        acc = (acc * x + v) % p
    return acc


def _interp2(s, y0, y1, alpha, p):
    """the line through (s, y0), (-s, y1) at alpha"""
    return (y0 + (y0 - y1) * (alpha - s) % p * pow(2 * s % p, p - 2, p)) % p


def _domain_index(x, log_n, F):
    p = F.p
    w_inv = pow(omega(F, log_n), p - 2, p)
    e, cur = 0, x % p
    for i in range(log_n):
        if pow(cur, 1 << (log_n - 1 - i), p) != 1:
            e |= 1 << i
            cur = cur * pow(w_inv, 1 << i, p) % p
    return e


def fri_verify_eval(field, hash_id, fri, proof, commitments, theta, poly_ids, combined_U, denominators, transcript):
    """zk::algorithms::verify_eval<FRI> (basic_fri.hpp:932-1150).  fri: lpc.FriParams; denominators: the points z of
    V = X - z; transcript: transcript.FiatShamirSequential (or anything with __call__, challenge, int_challenge)."""
    F = FIELD_BY_NAME[field] if isinstance(field, str) else field
    p, h, steps, log_d0 = F.p, HASH_BY_ID[hash_id], fri.step_list, fri.log_d0
    nbytes = (p.bit_length() + 7) // 8
    if not steps or steps[-1] != 1 or any(s < 1 or s > 10 for s in steps):
        return False
    try:
        return _fri_verify_eval(F, h, nbytes, fri, proof, commitments, theta, poly_ids, combined_U, denominators, transcript)
    except (KeyError, IndexError, TypeError, ValueError, AttributeError):
        return False         # malformed proof structure: reject, never raise


def _fri_verify_eval(F, h, nbytes, fri, proof, commitments, theta, poly_ids, combined_U, denominators, transcript):
    p, steps, log_d0 = F.p, fri.step_list, fri.log_d0
    if len(proof["fri_roots"]) != len(steps) or len(proof["query_proofs"]) != fri.lambda_:
        return False
    for qp in proof["query_proofs"]:
        if len(qp["round_proofs"]) != len(steps) or set(qp["initial_proof"]) != set(commitments):
            return False
    fp = proof["final_polynomial"]
    deg = max((i for i, v in enumerate(fp) if v), default=0)
    if deg > 2 ** (fri.degree_log - fri.r + 1) - 1:
        return False
    alphas = []
    for i, step in enumerate(steps):
        transcript(proof["fri_roots"][i])
        alphas += [transcript.challenge(p) for _ in range(step)]
    if fri.use_grinding:
        transcript(int(proof["proof_of_work"]).to_bytes(4, "big"))
        if transcript.int_challenge(32) & fri.grinding_parameter:
            return False
    sizes = lambda t: 1 << (log_d0 - t)     # noqa: E731  |D[t]|
    for q in range(fri.lambda_):
        qp = proof["query_proofs"][q]
        domain_size = sizes(0)
        x = pow(transcript.challenge(p), (p - 1) // domain_size, p)
        x_index = _domain_index(x, log_d0, F)
        s, s_idx = _calculate_s(x_index, steps[0], log_d0, F)
        order = _correct_order(x_index, domain_size, steps[0], s_idx)
        half = (1 << steps[0]) // 2
        for k, ip in qp["initial_proof"].items():
            if ip["p"]["root"] != commitments[k]:
                return False
            data = [v for vals in ip["values"] for pos in order for v in vals[pos]]
            if not _merkle_validate(ip["p"], _leaf(data, nbytes), h, _folded(x_index, domain_size, steps[0]), log_d0 - steps[0]):
                return False
        theta_acc = 1
        y = [[0, 0] for _ in range(half)]
        for pi in range(len(poly_ids)):
            Q = [[0, 0] for _ in range(half)]
            for (bk, bi) in poly_ids[pi]:
                vals = qp["initial_proof"][bk]["values"][bi]
                for j in range(half):
                    Q[j][0] = (Q[j][0] + vals[j][0] * theta_acc) % p
                    Q[j][1] = (Q[j][1] + vals[j][1] * theta_acc) % p
                theta_acc = theta_acc * theta % p
            for j in range(half):
                id0 = 0 if s_idx[j][0] < s_idx[j][1] else 1
                for side, sid in ((0, id0), (1, 1 - id0)):
                    v = (Q[j][side] - combined_U[pi]) * pow((s[j][sid] - denominators[pi]) % p, p - 2, p) % p
                    y[j][side] = (y[j][side] + v) % p
        t = 0
        for i, step in enumerate(steps):
            rp = qp["round_proofs"][i]
            if rp["p"]["root"] != proof["fri_roots"][i]:
                return False
            s, s_idx = _calculate_s(x_index, step, log_d0 - t, F)
            order = _correct_order(x_index, domain_size, step, s_idx)
            if not _merkle_validate(rp["p"], _leaf([v for pos in order for v in y[pos]], nbytes), h,
                                    _folded(x_index, domain_size, step), log_d0 - t - step):
                return False
            for _ in range(step - 1):            # colinear checks inside a multi-step round
                domain_size = sizes(t)
                x_index %= domain_size
                _, idx_next = _calculate_s(x_index % sizes(t + 1), step, log_d0 - t - 1, F)
                s, s_idx = _calculate_s(x_index, step, log_d0 - t, F)
                y_next = []
                for yi in range(len(y) // 2):
                    i0 = 0 if s_idx[2 * yi][0] < s_idx[2 * yi][1] else 1
                    left = _interp2(s[2 * yi][i0], y[2 * yi][0], y[2 * yi][1], alphas[t], p)
                    i0 = 0 if s_idx[2 * yi + 1][0] < s_idx[2 * yi + 1][1] else 1
                    right = _interp2(s[2 * yi + 1][i0], y[2 * yi + 1][0], y[2 * yi + 1][1], alphas[t], p)
                    y_next.append([left, right] if idx_next[yi][0] < idx_next[yi][1] else [right, left])
                y = y_next
                t += 1
            domain_size = sizes(t)
            x_index %= domain_size
            s, s_idx = _calculate_s(x_index, step, log_d0 - t, F)
            i0 = 0 if s_idx[0][0] < s_idx[0][1] else 1
            interpolant = _interp2(s[0][i0], y[0][0], y[0][1], alphas[t], p)
            ind = 0 if s_idx[0][i0] % (domain_size // 2) < domain_size // 4 else 1
            if interpolant != rp["y"][0][ind]:
                return False
            y = [list(v) for v in rp["y"]]
            if i < len(steps) - 1:
                t += 1
                domain_size = sizes(t)
                x_index %= domain_size
        x_index %= sizes(t)
        x = pow(omega(F, log_d0 - t), x_index, p)
        x = x * x % p
        ind = 0 if x_index % (sizes(t) // 2) < sizes(t) // 4 else 1
        if y[0][ind] != _horner(fp, x, p) or y[0][1 - ind] != _horner(fp, (p - x) % p, p):
            return False
    return True


def lpc_verify_eval(field, hash_id, fri, proof, points, commitments, transcript, fixed_batches=(), etha=None, fixed_values=None):
    """lpc_commitment_scheme::verify_eval (lpc.hpp:202-263).  points: {k: [[point, ..] per polynomial]}."""
    F = FIELD_BY_NAME[field] if isinstance(field, str) else field
    try:
        p, z = F.p, proof["z"]
        if set(z) != set(points) or set(z) != set(commitments):
            return False
        for k in z:
            if len(z[k]) != len(points[k]) or any(len(z[k][j]) != len(points[k][j]) for j in range(len(z[k]))):
                return False
        if fixed_batches and any(len(fixed_values[k]) != len(z[k]) for k in fixed_batches if k in z):
            return False
    except (KeyError, IndexError, TypeError, AttributeError):
        return False
    for k in sorted(commitments):
        transcript(commitments[k])
    uniq = []
    for k in sorted(points):
        for per_poly in points[k]:
            for x in per_poly:
                if x not in uniq:
                    uniq.append(x)
    total = len(uniq) + (1 if fixed_batches else 0)
    U, V, poly_map = [0] * total, [0] * total, [[] for _ in range(total)]
    theta = transcript.challenge(p)
    theta_acc = 1
    for pi, point in enumerate(uniq):
        V[pi] = point
        for k in sorted(z):
            for j in range(len(z[k])):
                if point in points[k][j]:
                    U[pi] = (U[pi] + z[k][j][points[k][j].index(point)] * theta_acc) % p
                    poly_map[pi].append((k, j))
                    theta_acc = theta_acc * theta % p
    if fixed_batches:
        pi = len(uniq)
        V[pi] = etha
        for k in sorted(z):
            if k in fixed_batches:
                for j in range(len(z[k])):
                    U[pi] = (U[pi] + fixed_values[k][j] * theta_acc) % p
                    poly_map[pi].append((k, j))
                    theta_acc = theta_acc * theta % p
    return fri_verify_eval(F, hash_id, fri, proof["fri_proof"], commitments, theta, poly_map, U, V, transcript)
