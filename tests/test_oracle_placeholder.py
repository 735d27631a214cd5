"""Pins oracle/placeholder.py by the permutation argument's own relation (permutation_argument.hpp:104-133): when the
column values respect the copy constraints of sigma, the grand product closes after the last row.  CPU only."""
import random

from oracle import fields, placeholder


def test_grand_product_closes_for_a_valid_permutation():
    F = fields.PALLAS_FP
    p, n, ncols = F.p, 32, 3
    rnd = random.Random(4)
    w = F.omega(5)
    deltas = [pow(F.g, i, p) for i in range(ncols)]                       # coset representatives, as the preprocessor uses
    ids = [[deltas[i] * pow(w, j, p) % p for j in range(n)] for i in range(ncols)]
    cells = [(i, j) for i in range(ncols) for j in range(n)]
    perm = cells[:]
    rnd.shuffle(perm)                                                        # sigma: cell -> cell (one big cycle structure)
    sigma = dict(zip(cells, perm))
    # values constant along every cycle of sigma
    val, seen = {}, set()
    for c in cells:
        if c in seen:
            continue
        v, x = rnd.randrange(p), c
        while x not in seen:
            seen.add(x)
            val[x] = v
            x = sigma[x]
    columns = [[val[(i, j)] for j in range(n)] for i in range(ncols)]
    s_sigma = [[ids[sigma[(i, j)][0]][sigma[(i, j)][1]] for j in range(n)] for i in range(ncols)]
    beta, gamma = rnd.randrange(p), rnd.randrange(p)
    V = placeholder.permutation_grand_product(columns, ids, s_sigma, beta, gamma, F)
    assert V[0] == 1 and len(V) == n
    nom = denom = 1
    for i in range(ncols):
        nom = nom * (columns[i][n - 1] + beta * ids[i][n - 1] + gamma) % p
        denom = denom * (columns[i][n - 1] + beta * s_sigma[i][n - 1] + gamma) % p
    assert V[n - 1] * nom % p * pow(denom, p - 2, p) % p == 1
    # and it does not close when one cell breaks a copy constraint
    columns[1][7] = (columns[1][7] + 1) % p
    V = placeholder.permutation_grand_product(columns, ids, s_sigma, beta, gamma, F)
    nom = denom = 1
    for i in range(ncols):
        nom = nom * (columns[i][n - 1] + beta * ids[i][n - 1] + gamma) % p
        denom = denom * (columns[i][n - 1] + beta * s_sigma[i][n - 1] + gamma) % p
    assert V[n - 1] * nom % p * pow(denom, p - 2, p) % p != 1


def test_scan_helpers():
    p = fields.BN254_FR.p
    x = [3, 5, 7, 11]
    assert placeholder.prefix_product(x, p) == [1, 3, 15, 105]
    assert placeholder.prefix_product(x, p, exclusive=False) == [3, 15, 105, 1155]
    assert [a * b % p for a, b in zip(x, placeholder.batch_inverse(x, p))] == [1] * 4


# ---------------------------------------------------------------------------------------------------------------------
# expressions, quotient, lookup sort: pinned on the relations the arguments prove
def chain_circuit_host(F, log_n, usable, seed):
    """host twin of crypto3_zk_b200.workloads.placeholder_chain_circuit with one (a, b, c) triple: a*b = c on the usable rows,
    a[j+1] = c[j] as copy constraints; returns the columns a, b, c, selector, q_last, q_blind, L_0, S_id, S_sigma"""
    p, n = F.p, 1 << log_n
    rnd = random.Random(seed)
    a, b, c = ([rnd.randrange(p) for _ in range(n)] for _ in range(3))
    for j in range(usable):
        if j:
            a[j] = c[j - 1]
        c[j] = a[j] * b[j] % p
    sel = [1 if j < usable else 0 for j in range(n)]
    q_last = [1 if j == usable else 0 for j in range(n)]
    q_blind = [1 if j > usable else 0 for j in range(n)]
    l0 = [1] + [0] * (n - 1)
    w = F.omega(log_n)
    s_id = [[pow(F.g, i, p) * pow(w, j, p) % p for j in range(n)] for i in range(3)]
    s_sigma = [row[:] for row in s_id]
    for j in range(usable - 1):
        s_sigma[0][j + 1], s_sigma[2][j] = s_id[2][j], s_id[0][j + 1]
    return a, b, c, sel, q_last, q_blind, l0, s_id, s_sigma


def test_expression_quotient_of_a_satisfied_circuit_is_exact():
    """gates_argument.hpp:133-217 + permutation_argument.hpp:170-215 + prover.hpp:260-283 on a satisfied circuit: every F_i
    vanishes on the basic domain, so the division by Z = X^n - 1 leaves no remainder and T Z = F; a broken gate leaves one"""
    from oracle import ntt
    F = fields.PALLAS_FP
    p, log_n = F.p, 3
    n, usable = 1 << log_n, 5
    a, b, c, sel, q_last, q_blind, l0, s_id, s_sigma = chain_circuit_host(F, log_n, usable, 3)
    beta, gamma, theta = 0x1234567, 0x7654321, 0xabcdef
    V = placeholder.permutation_grand_product([a, b, c], s_id, s_sigma, beta, gamma, F)
    assert V[usable] == 1
    columns = [a, b, c, sel, q_last, q_blind, l0, V] + s_id + s_sigma     # indices 0..7, S_id 8..10, S_sigma 11..13
    col = lambda i, r=0: ("col", i, r)
    one = ("const", 1)
    g = h = None
    for i in range(3):
        gi = ("add", ("add", ("mul", ("const", beta), col(8 + i)), ("const", gamma)), col(i))
        hi = ("add", ("add", ("mul", ("const", beta), col(11 + i)), ("const", gamma)), col(i))
        g = gi if g is None else ("mul", g, gi)
        h = hi if h is None else ("mul", h, hi)
    mask = ("sub", ("sub", one, col(4)), col(5))
    f = [("mul", ("sub", one, col(7)), col(6)),
         ("mul", mask, ("sub", ("mul", col(7, 1), h), ("mul", col(7), g))),
         ("mul", col(4), ("sub", ("mul", col(7), col(7)), col(7))),
         ("mul", ("mul", ("mul", ("sub", ("mul", col(0), col(1)), col(2)), ("const", theta)), col(3)), mask)]
    alphas = [11, 22, 33, 44]
    total = None
    for fi, al in zip(f, alphas):
        term = ("mul", fi, ("const", al))
        total = term if total is None else ("add", total, term)
    coeffs = placeholder.expr_polynomial(total, columns, F)
    # expr_dfs is the same polynomial in evaluation form: check at a few points of the extended subgroup by direct evaluation
    ext = 8 * n
    dfs = placeholder.expr_dfs(total, columns, F, ext)
    w_ext = F.omega(log_n + 3)
    for k in (0, 1, 9, ext - 1):
        x = pow(w_ext, k, p)
        assert dfs[k] == sum(cf * pow(x, i, p) for i, cf in enumerate(coeffs)) % p
    # on the basic domain (every 8th point) the expression is what the columns say row by row, and it is zero
    for j in range(n):
        val = placeholder.expr_at_point(total, lambda cidx, r: columns[cidx][(j + r) % n], p)
        assert dfs[8 * j] == val == 0
    chunks, rem = placeholder.quotient_split(coeffs, n, 6, F)
    assert rem == [0] * n
    t = [v for ch in chunks for v in ntt.dfs_coefficients(ch, F)]
    tz = [0] * (len(t) + n)
    for i, v in enumerate(t):                     # T (X^n - 1)
        tz[i + n] = (tz[i + n] + v) % p
        tz[i] = (tz[i] - v) % p
    while len(tz) > len(coeffs) and tz[-1] == 0:
        tz.pop()
    assert tz == coeffs + [0] * (len(tz) - len(coeffs))
    # a violated gate: the remainder is not zero
    c[2] = (c[2] + 1) % p
    columns[2] = c
    _, rem = placeholder.quotient_split(placeholder.expr_polynomial(f[3], columns, F), n, 6, F)
    assert any(rem)


def test_sort_polynomials_properties():
    """lookup_argument.hpp:565-633: the sorted columns are a permutation of table + inputs arranged in table order, which is
    what makes compute_V_L's product close (h over sorted equals g over input/table, :375-409)"""
    F = fields.BLS12_381_FR
    p = F.p
    rnd = random.Random(8)
    n, usable = 32, 27
    table_vals = [rnd.randrange(1, p) for _ in range(12)]
    value = [0] * 3 + [v for v in table_vals for _ in range(2)]        # leading zeros (padding), every value twice, adjacent
    value = (value + [0] * n)[:n]
    inputs = [[rnd.choice(table_vals + [0]) for _ in range(n)] for _ in range(2)]
    s = placeholder.sort_polynomials(inputs, [value], n, usable)
    assert len(s) == 3 and all(len(c) == n for c in s)
    flat = [v for c in s for v in c[:usable]]
    want = sorted(value[:usable] + inputs[0][:usable] + inputs[1][:usable])
    nz = [v for v in flat if v]
    assert sorted(nz) == [v for v in want if v]                          # every non-zero value as often as it occurs
    order = {v: i for i, v in enumerate(table_vals)}
    ranks = [order[v] for v in nz]
    assert ranks == sorted(ranks)                                        # in table order
    assert s[0][usable] == s[1][0] and s[1][usable] == s[2][0]
    # with the sorted columns the lookup grand product closes at the last usable row
    beta, gamma = rnd.randrange(p), rnd.randrange(p)
    V = placeholder.lookup_grand_product(inputs, [value], s, beta, gamma, usable, F)
    assert V[0] == 1 and V[usable] == 1
    inputs[1][4] = (inputs[1][4] + 1) % p              # an input outside the table: the sorted columns no longer balance it
    assert placeholder.lookup_grand_product(inputs, [value], s, beta, gamma, usable, F)[usable] != 1


def test_lookup_argument_relations_hold_row_by_row():
    """lookup_argument.hpp:153-325 on a two-column table (t, t^2) and one lookup constraint (u, u^2): with the compressed
    values / inputs of prepare_lookup_value / prepare_lookup_input, the sorted columns of sort_polynomials and V_L of
    compute_V_L, the four argument polynomials F_3 .. F_6 vanish on every row - when the table leaves row 0 zero, as
    upstream's table packing does (the sort starts its walk from a zero, :596); a table that starts in row 0 breaks V_L"""
    F = fields.PALLAS_FP
    p, n = F.p, 32
    usable, K = n - 4, 13
    rnd = random.Random(1)

    def relations(first_row):
        rows = range(first_row, first_row + K)
        u = [rnd.randrange(1, K + 1) for _ in range(usable)] + [rnd.randrange(p) for _ in range(n - usable)]
        v = [x * x % p for x in u[:usable]] + [rnd.randrange(p) for _ in range(n - usable)]
        c0 = [j - first_row + 1 if j in rows else 0 for j in range(n)]
        c1 = [x * x for x in c0]
        tag = [1 if j in rows else 0 for j in range(n)]
        lsel = [1 if j < usable else 0 for j in range(n)]
        q_last = [1 if j == usable else 0 for j in range(n)]
        q_blind = [1 if j > usable else 0 for j in range(n)]
        theta, beta, gamma = rnd.randrange(p), rnd.randrange(p), rnd.randrange(p)
        mask = [(1 - q_last[j] - q_blind[j]) % p for j in range(n)]
        value = [mask[j] * tag[j] * (1 + theta * c0[j] + theta * theta * c1[j]) % p for j in range(n)]
        inp = [lsel[j] * (1 + theta * u[j] + theta * theta * v[j]) % p for j in range(n)]
        s = placeholder.sort_polynomials([inp], [value], n, usable)
        V = placeholder.lookup_grand_product([inp], [value], s, beta, gamma, usable, F)
        ob, part1 = (1 + beta) % p, (1 + beta) * gamma % p
        g = [ob * (gamma + inp[j]) % p * ((part1 + value[j] + beta * value[(j + 1) % n]) % p) % p for j in range(n)]
        h = [1] * n
        for col in s:
            h = [h[j] * ((part1 + col[j] + beta * col[(j + 1) % n]) % p) % p for j in range(n)]
        f3 = [(1 if j == 0 else 0) * (1 - V[j]) % p for j in range(n)]
        f4 = [q_last[j] * (V[j] * V[j] - V[j]) % p for j in range(n)]
        f5 = [(g[j] * V[j] - h[j] * V[(j + 1) % n]) % p * ((q_last[j] + q_blind[j] - 1) % p) % p for j in range(n)]
        f6 = [(1 if j == 0 else 0) * (s[1][j] - s[0][(j + usable) % n]) % p for j in range(n)]
        return V[usable] == 1, [any(f) for f in (f3, f4, f5, f6)]

    closes, nonzero = relations(first_row=1)
    assert closes and nonzero == [False] * 4
    closes, nonzero = relations(first_row=0)
    assert not closes and nonzero[1]
