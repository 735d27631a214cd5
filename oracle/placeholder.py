"""Placeholder argument builders (oracle; test infrastructure only): the grand products of the permutation and lookup
arguments, expressions over columns by exact polynomial arithmetic, the quotient division / split and the lookup sort.

Literal restatement of zk/snark/systems/plonk/placeholder/permutation_argument.hpp:104-133 on Python integers.
Parity status: unpinned by reference fixtures (the reference holds no concrete V_P); the pin is the argument's own
relation - for columns that satisfy the copy constraints of a permutation sigma the product closes, i.e.
V_P[n-1] * ratio[n-1] = 1 (tests/test_oracle_placeholder.py).
"""


def permutation_grand_product(columns, s_id, s_sigma, beta, gamma, field):
    """columns / s_id / s_sigma: lists (one per permuted column, in global_indices order) of n values."""
    p = field.p
    n = len(columns[0])
    g_v = [[(c + beta * a + gamma) % p for c, a in zip(col, sid)] for col, sid in zip(columns, s_id)]
    h_v = [[(c + beta * a + gamma) % p for c, a in zip(col, sg)] for col, sg in zip(columns, s_sigma)]
    V = [1] * n
    for j in range(1, n):
        nom, denom = 1, 1
        for i in range(len(columns)):
            nom = nom * g_v[i][j - 1] % p
            denom = denom * h_v[i][j - 1] % p
        V[j] = V[j - 1] * nom % p * pow(denom, p - 2, p) % p
    return V


def prefix_product(x, p, exclusive=True):
    out, acc = [], 1
    for v in x:
        if exclusive:
            out.append(acc)
        acc = acc * v % p
        if not exclusive:
            out.append(acc)
    return out


def batch_inverse(x, p):
    return [pow(v, p - 2, p) for v in x]


def lookup_grand_product(reduced_input, reduced_value, sorted_, beta, gamma, usable_rows, field):
    """compute_V_L, lookup_argument.hpp:375-409 (lists of n-value lists)"""
    p = field.p
    n = len(sorted_[0])
    V = [0] * n
    V[0] = 1
    part1 = (1 + beta) * gamma % p
    for k in range(1, usable_rows + 1):
        g_tmp = pow(1 + beta, len(reduced_input), p)
        for col in reduced_input:
            g_tmp = g_tmp * (gamma + col[k - 1]) % p
        for col in reduced_value:
            g_tmp = g_tmp * (part1 + col[k - 1] + beta * col[k]) % p
        h_tmp = 1
        for col in sorted_:
            h_tmp = h_tmp * (part1 + col[k - 1] + beta * col[k]) % p
        V[k] = V[k - 1] * g_tmp % p * pow(h_tmp, p - 2, p) % p
    return V


# ---------------------------------------------------------------------------------------------------------------------
# Expressions over columns, quotient, lookup sort.  Expressions are nested tuples
#   ("col", c, rot) | ("const", v) | ("add", a, b) | ("sub", a, b) | ("mul", a, b) | ("neg", a)
# and are evaluated here by EXACT polynomial arithmetic in coefficient form (naive products), independently of the
# device's point-by-point evaluation: gates_argument.hpp:76-217 and permutation_argument.hpp:170-215 build the same
# polynomials through polynomial_dfs operators (which resize to the product's degree and multiply pointwise).

def _poly_add(a, b, p):
    n = max(len(a), len(b))
    return [((a[i] if i < len(a) else 0) + (b[i] if i < len(b) else 0)) % p for i in range(n)]


def _poly_mul(a, b, p):
    out = [0] * (len(a) + len(b) - 1)
    for i, x in enumerate(a):
        if x:
            for j, y in enumerate(b):
                out[i + j] = (out[i + j] + x * y) % p
    return out


def expr_polynomial(expr, columns, field):
    """Coefficients of the polynomial the expression denotes.  columns: lists of n evaluations on the basic domain;
    a rotation r is the shift f(omega^r X) (math::polynomial_shift, gates_argument.hpp:113-115)."""
    from . import ntt
    p = field.p
    kind = expr[0]
    if kind == "col":
        vals = columns[expr[1]]
        n = len(vals)
        c = ntt.dfs_coefficients(vals, field)
        w = pow(field.omega(n.bit_length() - 1), expr[2] % n, p)
        return [v * pow(w, k, p) % p for k, v in enumerate(c)]
    if kind == "const":
        return [expr[1] % p]
    if kind == "neg":
        return [(-v) % p for v in expr_polynomial(expr[1], columns, field)]
    a, b = expr_polynomial(expr[1], columns, field), expr_polynomial(expr[2], columns, field)
    if kind == "add":
        return _poly_add(a, b, p)
    if kind == "sub":
        return _poly_add(a, [(-v) % p for v in b], p)
    if kind == "mul":
        return _poly_mul(a, b, p)
    raise ValueError(kind)


def expr_dfs(expr, columns, field, ext_size):
    """evaluation form of the expression's polynomial on the subgroup of size ext_size (its degree must be below it)"""
    from . import ntt
    c = expr_polynomial(expr, columns, field)
    while len(c) > 1 and c[-1] == 0:
        c.pop()
    assert len(c) <= ext_size, "expression degree exceeds the extended domain"
    a = c + [0] * (ext_size - len(c))
    ntt.EvaluationDomain(field, ext_size).fft(a)
    return a


def quotient_split(f_coeffs, n, nchunks, field):
    """T = F / (X^n - 1) by long division (remainder dropped, placeholder/prover.hpp:268-283), split into chunks of n
    coefficients (detail::split_polynomial, :47-70), each chunk from_coefficients() on the basic domain (:255-257)."""
    from . import ntt
    p = field.p
    rem = list(f_coeffs)
    q = [0] * max(len(rem) - n, 0)
    for k in range(len(rem) - 1, n - 1, -1):       # eliminate the top coefficient with X^(k-n) (X^n - 1)
        t = rem[k]
        if t:
            q[k - n] = (q[k - n] + t) % p
            rem[k] = 0
            rem[k - n] = (rem[k - n] + t) % p
    out = []
    for ch in range(nchunks):
        c = q[ch * n:(ch + 1) * n]
        c = c + [0] * (n - len(c))
        ntt.EvaluationDomain(field, n).fft(c)
        out.append(c)
    return out, rem[:n]


def sort_polynomials(reduced_input, reduced_value, domain_size, usable_rows):
    """lookup_argument.hpp:565-633, line by line."""
    sorting_map = {}
    for col in reduced_value:
        for j in range(usable_rows):
            sorting_map[col[j]] = sorting_map.get(col[j], 0) + 1
    for col in reduced_input:
        for j in range(usable_rows):
            assert col[j] in sorting_map
            sorting_map[col[j]] += 1
    sorted_ = [[0] * domain_size for _ in range(len(reduced_input) + len(reduced_value))]
    pos = [0, 0]

    def append(v):
        sorted_[pos[0]][pos[1]] = v
        pos[1] += 1
        if pos[1] >= usable_rows:
            pos[0] += 1
            pos[1] = 0
    prev = 0
    for col in reduced_value:
        for j in range(usable_rows):
            if col[j] != prev:
                if prev == 0:
                    append(prev)
                else:
                    for _ in range(sorting_map[prev]):
                        append(prev)
                prev = col[j]
    if prev != 0:
        for _ in range(sorting_map[prev]):
            append(prev)
    for i in range(len(sorted_) - 1):
        sorted_[i][usable_rows] = sorted_[i + 1][0]
    return sorted_


def expr_at_point(expr, value_of, p):
    """The expression with every ("col", c, rot) replaced by value_of(c, rot): how the verifier evaluates the argument
    polynomials at the challenge from the opened column values (placeholder/verifier.hpp:233-275, gates_argument.hpp:219-262)."""
    kind = expr[0]
    if kind == "col":
        return value_of(expr[1], expr[2]) % p
    if kind == "const":
        return expr[1] % p
    if kind == "neg":
        return (-expr_at_point(expr[1], value_of, p)) % p
    a, b = expr_at_point(expr[1], value_of, p), expr_at_point(expr[2], value_of, p)
    if kind == "add":
        return (a + b) % p
    if kind == "sub":
        return (a - b) % p
    if kind == "mul":
        return a * b % p
    raise ValueError(kind)
