"""Opening side of the LPC scheme (oracle; test infrastructure only): eval_polys and the combined quotient Q.

In-repo, exact:
  polys_evaluator::eval_polys      zk/commitments/batched_commitment.hpp:176-190
  get_unique_points                zk/commitments/batched_commitment.hpp (first-seen order over batches/polys/points)
  lpc proof_eval: theta powers, Q  zk/commitments/polynomial/lpc.hpp:113-181
Polynomials are coefficient lists (math::polynomial) here; dfs inputs are converted with ntt.dfs_coefficients
exactly where the reference calls `.coefficients()`.
"""
from .ntt import dfs_coefficients, dfs_from_coefficients


def poly_eval(c, x, p):
    acc = 0
    for v in reversed(c):
        acc = (acc * x + v) % p
    return acc


def poly_add(a, b, p):
    n = max(len(a), len(b))
    return [((a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0)) % p for i in range(n)]


def poly_div_linear(c, z, p):
    """Q = c / (X - z): quotient of the polynomial division, remainder dropped (math::polynomial operator/)."""
    n = len(c)
    if n <= 1:
        return []
    q = [0] * (n - 1)
    acc = 0
    for k in range(n - 1, 0, -1):
        acc = (acc * z + c[k]) % p
        q[k - 1] = acc
    return q


def eval_polys(polys, points, field):
    """polys[k][i]: dfs value lists; points[k][i]: list of points -> z[k][i][j] (batched_commitment.hpp:176-190)."""
    p = field.p
    z = {}
    for k in polys:
        z[k] = []
        for i, poly in enumerate(polys[k]):
            co = dfs_coefficients(poly, field)
            z[k].append([poly_eval(co, x, p) for x in points[k][i]])
    return z


def unique_points(points):
    out = []
    for k in sorted(points):
        for per_poly in points[k]:
            for x in per_poly:
                if x not in out:
                    out.append(x)
    return out


def combined_q(polys, points, z, theta, field, fixed_batches=(), etha=None, fixed_values=None):
    """lpc.hpp:126-181: for every unique point, Q_point = sum_{(i,j) opened there} theta^acc (g_ij - z_ij) / (X - point)
    with theta_acc running across points in the reference's loop order; then the fixed batches at etha.
    Returns (combined_Q coefficient list, combined_Q as dfs of the next power of two = from_coefficients)."""
    p = field.p
    theta_acc = 1
    combined = []
    for point in unique_points(points):
        q = []
        for k in sorted(polys):
            for i, poly in enumerate(polys[k]):
                if point not in points[k][i]:
                    continue
                j = points[k][i].index(point)
                g = [v * theta_acc % p for v in dfs_coefficients(poly, field)]
                q = poly_add(q, g, p)
                q[0] = (q[0] - z[k][i][j] * theta_acc) % p
                theta_acc = theta_acc * theta % p
        combined = poly_add(combined, poly_div_linear(q, point, p), p)
    for k in sorted(polys):
        if k not in fixed_batches:
            continue
        q = []
        for i, poly in enumerate(polys[k]):
            g = [v * theta_acc % p for v in dfs_coefficients(poly, field)]
            q = poly_add(q, g, p)
            q[0] = (q[0] - fixed_values[k][i] * theta_acc) % p
            theta_acc = theta_acc * theta % p
        combined = poly_add(combined, poly_div_linear(q, etha, p), p)
    return combined, dfs_from_coefficients(combined, field)
