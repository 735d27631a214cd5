"""FRI/LPC commit-side restatement (oracle; test infrastructure only).

In-repo, exact:
  fold_polynomial (dfs form)     zk/commitments/detail/polynomial/fold_polynomial.hpp:68-93
  precommit (leaf index pattern) zk/commitments/detail/polynomial/basic_fri.hpp:349-351,445-496
  leaf serialisation             zk/detail/field_element_consumer.hpp:87-95 + basic_fri.hpp:96-100
                                 (big-endian canonical integer, ceil(modulus_bits/8) bytes)
  lpc commit = root of that tree zk/commitments/polynomial/lpc.hpp:101-106
Upstream (believed): make_merkle_tree<Hash,2>: leaf digest = H(leaf bytes),
inner = H(left || right), #leaves a power of two.  The reference holds no concrete root vector,
so Merkle-root parity is "unpinned" by reference fixtures.
"""
from .ntt import EvaluationDomain, dfs_resize


def fold_polynomial_dfs(f, alpha, field, domain_size=None):
    """f'[i] = 1/2 * ((1 + alpha*w^-i) f[i] + (1 - alpha*w^-i) f[i + |D|/2]),  i < |D|/2."""
    p = field.p
    n = len(f) if domain_size is None else domain_size
    dom = EvaluationDomain(field, n)
    two_inv = field.inv(2)
    omega_inv = dom.get_domain_element(n - 1)
    acc = alpha % p
    out = []
    for i in range(n // 2):
        out.append(two_inv * ((1 + acc) * f[i] + (1 - acc) * f[n // 2 + i]) % p)
        acc = acc * omega_inv % p
    return out


def fold_polynomial_coeffs(f, alpha, p):
    """Coefficient form (fold_polynomial.hpp:49-66): f'(X) = f_even(X) + alpha*f_odd(X)."""
    f = list(f)
    if len(f) % 2:
        f.append(0)
    return [(f[2 * i] + alpha * f[2 * i + 1]) % p for i in range(len(f) // 2)]


def leaf_indices(x_index, domain_size, fri_step):
    """Order in which one polynomial's evaluations enter leaf x_index (basic_fri.hpp:469-490, m=2)."""
    coset_size = 1 << fri_step
    s = [[0, 0] for _ in range(coset_size // 2)]
    s[0] = [x_index, (x_index + domain_size // 2) % domain_size]
    order = list(s[0])
    base_index = domain_size // 4
    prev_half = 1
    i = 1
    while i < coset_size // 2:
        for j in range(prev_half):
            a = (base_index + s[j][0]) % domain_size
            s[i] = [a, (a + domain_size // 2) % domain_size]
            order += s[i]
            i += 1
        base_index //= 2
        prev_half <<= 1
    return order


def leaf_bytes(polys, x_index, domain_size, fri_step, field):
    nb = field.nbytes
    out = bytearray()
    for poly in polys:                      # polynomial is the outer loop inside a leaf (:468)
        for idx in leaf_indices(x_index, domain_size, fri_step):
            out += int(poly[idx]).to_bytes(nb, "big")
    return bytes(out)


def merkle_tree(leaves, h):
    """Returns list of levels, level 0 = leaf digests; root = levels[-1][0]."""
    n = len(leaves)
    assert n >= 1 and n & (n - 1) == 0
    level = [h(l) for l in leaves]
    levels = [level]
    while len(level) > 1:
        level = [h(level[2 * i] + level[2 * i + 1]) for i in range(len(level) // 2)]
        levels.append(level)
    return levels


def merkle_proof(levels, leaf_idx):
    path = []
    for lvl in levels[:-1]:
        path.append(lvl[leaf_idx ^ 1])
        leaf_idx >>= 1
    return path


def precommit(polys_dfs, field, domain_size, fri_step, h):
    """precommit(container<polynomial_dfs>, D, fri_step): resize every poly to |D| then build the
    tree.  Returns (levels, resized polys)."""
    ext = [p if len(p) == domain_size else dfs_resize(p, field, domain_size) for p in polys_dfs]
    leafs_number = domain_size >> fri_step
    leaves = [leaf_bytes(ext, x, domain_size, fri_step, field) for x in range(leafs_number)]
    return merkle_tree(leaves, h), ext


def lpc_commit(polys_dfs, field, degree_log, expand_factor, fri_step, h):
    """lpc_commitment_scheme::commit: D[0] has size 2^(degree_log+expand_factor) (basic_fri.hpp:162)."""
    levels, _ = precommit(polys_dfs, field, 1 << (degree_log + expand_factor), fri_step, h)
    return levels[-1][0]


def commit_phase(f, field, log_n, step_list, h, transcript):
    """Commit phase of zk::algorithms::proof_eval<FRI> (basic_fri.hpp:706-737), dfs form, literal order:
    for every round: push f, tree = precommit(f, D[t], step_list[i]) (round 0: combined_Q_precommitment),
    transcript(root), then step_list[i] times { alpha = transcript.challenge(); f = fold(f, alpha, D[t]); t++ }.
    `transcript`: object with absorb(bytes) and challenge(field) (oracle.hashes.FiatShamirSequential).
    Returns dict(fs, roots, alphas, final_polynomial, levels)."""
    from .ntt import dfs_coefficients
    f = list(f)
    fs, roots, alphas, all_levels = [], [], [], []
    t = 0
    for i, step in enumerate(step_list):
        fs.append(f)
        levels, _ = precommit([f], field, 1 << (log_n - t), step, h)
        all_levels.append(levels)
        roots.append(levels[-1][0])
        transcript.absorb(levels[-1][0])
        for _ in range(step):
            alphas.append(transcript.challenge(field))
            f = fold_polynomial_dfs(f, alphas[t], field, 1 << (log_n - t))
            t += 1
    fs.append(f)
    return {"fs": fs, "roots": roots, "alphas": alphas, "final_polynomial": dfs_coefficients(f, field), "levels": all_levels}
