"""Placeholder argument builders (oracle; test infrastructure only): the grand product of the permutation argument.

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
