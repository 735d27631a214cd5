"""Synthetic workloads of BASELINE.json's configs #4 and #5 at full size; used by bench.py and by the
size-independent GPU tests.  Nothing here is a compute path: the builders only lay out inputs.

  groth16_field_input_example   generate_r1cs_example_with_field_input (test/systems/ppzksnark/r1cs_examples.hpp:77-146)
                                as CSR arrays built with numpy, plus its satisfying assignment
  groth16_synthetic_key         query vectors of the right sizes made of distinct valid curve points (P_i = A[i % m] + B[i / m])
  placeholder_batches           column shapes of the widest circuit of the reference's Placeholder tests
                                (test/systems/plonk/placeholder/circuits.hpp:575-579: 30 witness + 1 public + 1 selector)
"""
import numpy as np

from .api import _int_rows
from .fields import CURVE_BY_NAME, FIELD_BY_NAME, coord_limbs
from .groth16 import ProvingKey, R1csConstraintSystem


def _csr(rows_cols, nrows):
    """rows_cols: list of (row array, col array) pieces -> (row_ptr, cols, ones) sorted by row (stable)."""
    r = np.concatenate([a for a, _ in rows_cols]).astype(np.int64)
    c = np.concatenate([b for _, b in rows_cols]).astype(np.uint32)
    order = np.argsort(r, kind="stable")
    r, c = r[order], c[order]
    row_ptr = np.zeros(nrows + 1, dtype=np.uint64)
    np.add.at(row_ptr, r + 1, 1)
    row_ptr = np.cumsum(row_ptr).astype(np.uint64)
    vals = np.zeros((c.size, 8), dtype=np.uint32)
    vals[:, 0] = 1
    return row_ptr, c, vals


def groth16_field_input_example(field, num_constraints, num_inputs, seed=0):
    """Returns (cs_shape, csr_sides, full_assignment_limbs): the chain a+b=c / a*b=c alternating, closed by
    (sum x)(sum x) = fin^2, after swap_AB_if_beneficial (r1cs.hpp:193-218).  cs_shape is an R1csConstraintSystem
    without term lists (sizes only); csr_sides = [A, B, C] as (row_ptr, col, values)."""
    import random
    F = FIELD_BY_NAME[field] if isinstance(field, str) else field
    p = F.p
    nc = num_constraints
    nv = nc + 2
    rnd = random.Random(seed)
    a, b = rnd.randrange(p), rnd.randrange(p)
    full = [a, b]
    for i in range(nc - 1):
        tmp = a * b % p if i % 2 else (a + b) % p
        full.append(tmp)
        a, b = b, tmp
    fin = sum(full[:nv - 1]) % p
    full.append(fin * fin % p)
    i = np.arange(nc - 1, dtype=np.int64)
    odd, even = i[i % 2 == 1], i[i % 2 == 0]
    last = np.full(nv - 1, nc - 1, dtype=np.int64)
    allv = np.arange(1, nv, dtype=np.int64)
    A = [(odd, odd + 1), (even, even + 1), (even, even + 2), (last, allv)]
    B = [(odd, odd + 2), (even, np.zeros_like(even)), (last, allv)]
    C = [(i, i + 3), (np.array([nc - 1]), np.array([nv]))]
    touched = lambda side: np.unique(np.concatenate([c for _, c in side])).size
    if touched(B) > touched(A):
        A, B = B, A
    sides = [_csr(s, nc) for s in (A, B, C)]
    cs = R1csConstraintSystem(num_inputs, nv - num_inputs, [None] * nc)
    return cs, sides, _int_rows([1] + full)


def curve_grid_points(ctx, curve, n, seed=7):
    """n distinct valid points of `curve` on the device: P_i = A[i % 1024] + B[i // 1024] with A, B = k*G tables
    computed through the product's own MSM entry point."""
    C = CURVE_BY_NAME[curve] if isinstance(curve, str) else curve
    cl = coord_limbs(C)
    gen = np.zeros((1, 2, cl), dtype=np.uint32)
    for k, coord in enumerate((C.gen_x, C.gen_y)):
        parts = coord if C.deg == 2 else (coord,)
        h = cl // len(parts)
        for j, v in enumerate(parts):
            for l in range(h):
                gen[0, k, j * h + l] = (int(v) >> (32 * l)) & 0xFFFFFFFF
    gb = ctx.msm_bases(C.name, gen)
    m = 1024 if n >= 1024 else max(1, n)
    nbt = (n + m - 1) // m
    rng = np.random.Generator(np.random.PCG64(seed))
    ks = rng.integers(0, 1 << 32, size=(m + nbt, 8), dtype=np.uint64).astype(np.uint32)
    ks[:, 7] &= 0x0FFFFFFF
    tabs = np.zeros((m + nbt, 2, cl), dtype=np.uint32)
    res = np.zeros(2 * cl, dtype=np.uint32)
    import ctypes
    from . import capi
    for i in range(m + nbt):
        capi.check(capi.lib().zkb_msm(ctx._h, gb._h, 0, 1, ks[i:i + 1].ctypes.data, capi.MEM_HOST, capi.u32_ptr(res),
                                      ctypes.c_void_p(0)), ctx._h)
        tabs[i] = res.reshape(2, cl)
    gb.free()
    return ctx.grid_points(C.name, n, tabs[:m], tabs[m:])


def groth16_synthetic_key(ctx, curve_g1, curve_g2, cs, csr_sides, precompute=False):
    """A proving key with query vectors of the reference's sizes (generator.hpp:176-205) filled with synthetic points:
    A_query nv+1, B_query dense (every variable occurs in B for the field-input example), H_query m-1, L_query nv-ni."""
    nv, ni = cs.num_variables, cs.num_inputs
    m = cs.num_constraints + ni + 1
    n1 = max(nv + 1, m)
    g1 = curve_grid_points(ctx, curve_g1, n1 + 4, seed=3)
    g2 = curve_grid_points(ctx, curve_g2, nv + 3, seed=4)
    G1, G2 = CURVE_BY_NAME[curve_g1], CURVE_BY_NAME[curve_g2]

    def pt(t, i, curve):
        from .api import _affine_from_limbs
        return _affine_from_limbs(t[i].cpu().numpy().view(np.uint32).reshape(-1), coord_limbs(curve), curve.deg)

    return ProvingKey(ctx, curve_g1, curve_g2, cs, pt(g1, n1, G1), pt(g1, n1 + 1, G1), pt(g2, nv + 1, G2), pt(g1, n1 + 2, G1),
                      pt(g2, nv + 2, G2), g1[:nv + 1], list(range(nv + 1)), g2[:nv + 1], g1[1:nv + 2], g1[:m - 1],
                      g1[2:nv - ni + 2], csr=csr_sides, precompute=precompute)


# Placeholder commitment phase (config #5): batches of lpc_commitment_scheme as the prover fills them
# (zk/snark/systems/plonk/placeholder/prover.hpp:141,170,202,213; preprocessor.hpp:481-489), widest test circuit.
PLACEHOLDER_CIRCUIT5 = {"witness": 30, "public": 1, "constant": 0, "selector": 1}


def placeholder_batches(shape=PLACEHOLDER_CIRCUIT5, quotient_chunks=4):
    """polynomial counts per batch: fixed = identity + sigma permutation polynomials of every permuted column,
    the two special selectors (q_last, q_blind), constants and selectors; variable = witness + public;
    permutation = V_P; quotient = T split in chunks."""
    permuted = shape["witness"] + shape["public"] + shape["constant"]
    return {0: 2 * permuted + 2 + shape["constant"] + shape["selector"],
            1: shape["witness"] + shape["public"], 2: 1, 3: quotient_chunks}


def _const_column(torch, value, n, device):
    row = torch.from_numpy(_int_rows([int(value)]).view(np.int32)).to(device)
    return row.expand(n, 8).contiguous()


def _rand_column(torch, shape, seed, device):
    g = torch.Generator(device=device).manual_seed(seed)
    x = torch.randint(-2**31, 2**31 - 1, shape, dtype=torch.int32, device=device, generator=g)
    x[..., 7] &= 0x0FFFFFFF      # uniform 252-bit values: canonical in every field of the table
    return x


def placeholder_chain_circuit(ctx, field, log_n, triples=1, usable_rows=None, seed=1, device="cuda", rotation_gate=True,
                              max_quotient_chunks=0, lookup=False):
    """A satisfiable synthetic Placeholder circuit with its assignment, all columns built on the device.

    Witness columns (a_k, b_k, c_k) for k < triples, one public-input column, one selector.  On every usable row the gate
    a_k * b_k - c_k = 0 holds, and (with rotation_gate) a_k(next row) - c_k = 0 on the usable rows but the last; the copy
    constraints a_k[j+1] = c_k[j] tie the same cells through the permutation argument, and public[0] = a_0[0].  Rows from
    `usable_rows` on hold random blinding values (q_last at usable_rows, q_blind after it: preprocessor.hpp:463-476).
    Identity / sigma polynomials as preprocessor.hpp:418-460 builds them (S_id[i][j] = delta^i omega^j, delta = the
    multiplicative generator).  With `lookup`: two more witness columns (u, v = u^2), a two-column lookup table
    (t, t^2), t = 1 .. K, in two constant columns with its tag selector, and a lookup gate (u, v) in table 1 on every
    usable row.  Returns (PlaceholderCircuit, witness [count, n, 8], public_input [1, n, 8])."""
    import torch
    from . import capi
    from .fields import omega
    from .placeholder import PlaceholderCircuit, col, mul, sub
    F = FIELD_BY_NAME[field] if isinstance(field, str) else field
    n = 1 << log_n
    usable = n - 4 if usable_rows is None else usable_rows
    assert 2 <= usable < n
    nw = 3 * triples + (2 if lookup else 0)
    witness = _rand_column(torch, (nw, n, 8), seed, device)
    for k in range(triples):
        b = witness[3 * k + 1]
        excl = ctx.prefix_product(F.name, b)                                  # [1, b0, b0 b1, ..]
        a = ctx.vec(F.name, capi.VEC_MUL, excl, witness[3 * k, 0].expand(n, 8).contiguous())
        c = ctx.vec(F.name, capi.VEC_MUL, a, b)
        witness[3 * k, :usable] = a[:usable]
        witness[3 * k + 2, :usable] = c[:usable]
    public = torch.zeros((1, n, 8), dtype=torch.int32, device=device)
    public[0, 0] = witness[0, 0]
    # selectors
    n_gate_sel = 2 if rotation_gate else 1
    selector = torch.zeros((n_gate_sel + (2 if lookup else 0), n, 8), dtype=torch.int32, device=device)
    selector[0, :usable, 0] = 1
    if rotation_gate:
        selector[1, :usable - 1, 0] = 1
    constants, lookup_tables, lookup_gates = None, None, None
    if lookup:
        K = max(2, min(usable // 2 - 1, 1 << 15))
        t = torch.arange(1, K + 1, dtype=torch.int32, device=device)
        constants = torch.zeros((2, n, 8), dtype=torch.int32, device=device)
        # the table occupies rows 1 .. K: row 0 stays zero, as upstream's table packing leaves it - sort_polynomials starts
        # its walk from a zero (lookup_argument.hpp:596), so the pair (0, first value) must exist in the table column too
        constants[0, 1:K + 1, 0] = t
        constants[1, 1:K + 1, 0] = t * t
        selector[n_gate_sel, 1:K + 1, 0] = 1                                 # the table's tag
        selector[n_gate_sel + 1, :usable, 0] = 1                             # the lookup gate's selector
        g = torch.Generator(device=device).manual_seed(seed + 99)
        u = torch.randint(1, K + 1, (usable,), dtype=torch.int32, device=device, generator=g)
        cu, cv = 3 * triples, 3 * triples + 1
        witness[cu, :usable] = 0
        witness[cv, :usable] = 0
        witness[cu, :usable, 0] = u
        witness[cv, :usable, 0] = u * u
        lookup_tables = [(n_gate_sel, [[0, 1]])]
        lookup_gates = [(n_gate_sel + 1, [(1, [col(cu), col(cv)])])]
    q_last = torch.zeros((n, 8), dtype=torch.int32, device=device)
    q_last[usable, 0] = 1
    q_blind = torch.zeros((n, 8), dtype=torch.int32, device=device)
    q_blind[usable + 1:, 0] = 1
    lagrange_0 = torch.zeros((n, 8), dtype=torch.int32, device=device)
    lagrange_0[0, 0] = 1
    # permutation polynomials
    npc = nw + 1
    powers = ctx.prefix_product(F.name, _const_column(torch, omega(F, log_n), n, device))
    s_id = torch.empty((npc, n, 8), dtype=torch.int32, device=device)
    for i in range(npc):
        s_id[i] = ctx.vec(F.name, capi.VEC_MUL, powers, _const_column(torch, pow(F.generator, i, F.p), n, device))
    s_sigma = s_id.clone()
    for k in range(triples):
        s_sigma[3 * k, 1:usable] = s_id[3 * k + 2, 0:usable - 1]
        s_sigma[3 * k + 2, 0:usable - 1] = s_id[3 * k, 1:usable]
    s_sigma[nw, 0] = s_id[0, 0]
    s_sigma[0, 0] = s_id[nw, 0]
    gates = [(0, [sub(mul(col(3 * k), col(3 * k + 1)), col(3 * k + 2)) for k in range(triples)])]
    if rotation_gate:
        gates.append((1, [sub(col(3 * k, 1), col(3 * k + 2)) for k in range(triples)]))
    circuit = PlaceholderCircuit(F, log_n, nw, 1, 2 if lookup else 0, selector.shape[0], gates, list(range(npc)), s_id, s_sigma, q_last,
                                 q_blind, lagrange_0, constants, selector, usable, max_quotient_chunks=max_quotient_chunks,
                                 lookup_tables=lookup_tables, lookup_gates=lookup_gates)
    return circuit, witness, public
