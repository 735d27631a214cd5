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
