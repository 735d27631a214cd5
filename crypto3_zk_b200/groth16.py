"""Host mirror of r1cs_gg_ppzksnark_prover::process (zk/snark/systems/ppzksnark/r1cs_gg_ppzksnark/prover.hpp:73-158)
over device-resident keys and vectors.

Everything data-parallel runs in libzkb200.so:
  rows of the constraint system   zkb_sparse_matvec                 (r1cs_to_qap.hpp:239-248, 289-291)
  witness map                     zkb_ntt x 7, zkb_vec              (r1cs_to_qap.hpp:250-321; the three vectors are one batch)
  evaluation_At / Bt / Ht / Lt    zkb_msm on resident query vectors (prover.hpp:108-139)
The O(1) group operations of the proof assembly (prover.hpp:141-155) are folded into the MSMs as extra base points:
  A = alpha + sum x_i A_i + r delta           = MSM([alpha_g1, delta_g1] ++ A_query, [1, r] ++ x)
  B = beta + sum x_i B_i + s delta            (same, on G2 and on G1)
  C = Ht + Lt + s A + r B_g1 - r s delta_g1   = one 5-point MSM over the intermediate results
so no curve arithmetic runs on the host.  Points are affine (x, y) integers ((c0, c1) pairs on G2), None = infinity.
"""
import numpy as np

from . import capi
from .api import _int_rows
from .fields import CURVE_BY_NAME, FIELD_BY_NAME, coord_limbs


def _points_array(curve, pts):
    """list of affine points (None = infinity) -> [n, 2, coord_limbs] uint32"""
    cl = coord_limbs(curve)
    out = np.zeros((len(pts), 2, cl), dtype=np.uint32)
    for i, P in enumerate(pts):
        if P is None:
            continue
        for k in range(2):
            parts = P[k] if curve.deg == 2 else (P[k],)
            h = cl // len(parts)
            for j, v in enumerate(parts):
                for l in range(h):
                    out[i, k, j * h + l] = (int(v) >> (32 * l)) & 0xFFFFFFFF
    return out


class R1csConstraintSystem:
    """r1cs_constraint_system: constraints as (a, b, c) lists of (variable index, coefficient) terms, index 0 = 1."""

    def __init__(self, num_inputs, num_aux, constraints):
        self.num_inputs, self.num_aux, self.constraints = int(num_inputs), int(num_aux), constraints

    @property
    def num_variables(self):
        return self.num_inputs + self.num_aux

    @property
    def num_constraints(self):
        return len(self.constraints)

    def csr(self, side):
        """CSR arrays (row_ptr u64, col u32, values [nnz, 8] u32) of one side (0 = a, 1 = b, 2 = c)."""
        row_ptr = np.zeros(self.num_constraints + 1, dtype=np.uint64)
        cols, vals = [], []
        for i, con in enumerate(self.constraints):
            for idx, co in con[side]:
                cols.append(idx)
                vals.append(co)
            row_ptr[i + 1] = len(cols)
        return row_ptr, np.asarray(cols, dtype=np.uint32), _int_rows(vals)


class ProvingKey:
    """r1cs_gg_ppzksnark proving key resident on one GPU.  Query vectors are lists of affine points or already
    encoded [n, 2, limbs] arrays / device tensors."""

    def __init__(self, ctx, curve_g1, curve_g2, cs, alpha_g1, beta_g1, beta_g2, delta_g1, delta_g2, A_query, B_indices,
                 B_g2, B_g1, H_query, L_query, csr=None, precompute=False):
        import torch
        self.ctx = ctx
        self.g1 = CURVE_BY_NAME[curve_g1] if isinstance(curve_g1, str) else curve_g1
        self.g2 = CURVE_BY_NAME[curve_g2] if isinstance(curve_g2, str) else curve_g2
        self.F = FIELD_BY_NAME[self.g1.scalar_field]
        self.cs = cs
        self.delta_g1 = delta_g1
        dev = "cuda:%d" % ctx.device

        def enc(curve, head, query):
            h = _points_array(curve, head)
            if isinstance(query, list):
                return np.concatenate([h, _points_array(curve, query)])
            if isinstance(query, np.ndarray):
                return np.concatenate([h, query])
            return torch.cat([torch.from_numpy(h.view(np.int32)).to(query.device), query])

        self.A = ctx.msm_bases(self.g1.name, enc(self.g1, [alpha_g1, delta_g1], A_query))
        self.B2 = ctx.msm_bases(self.g2.name, enc(self.g2, [beta_g2, delta_g2], B_g2))
        self.B1 = ctx.msm_bases(self.g1.name, enc(self.g1, [beta_g1, delta_g1], B_g1))
        self.H = ctx.msm_bases(self.g1.name, enc(self.g1, [], H_query))
        self.L = ctx.msm_bases(self.g1.name, enc(self.g1, [], L_query))
        if precompute:
            for b in (self.A, self.B1, self.H, self.L):
                b.precompute()
        self.B_index = torch.as_tensor(np.asarray(B_indices, dtype=np.int64), device=dev)
        nc = cs.num_constraints
        m = nc + cs.num_inputs + 1
        if m & (m - 1):
            raise capi.ZkbInvalidArgument(capi.ERR_INVALID_ARGUMENT,
                                          "num_constraints + num_inputs + 1 must be a power of two (basic_radix2_domain)")
        self.m, self.log_m = m, m.bit_length() - 1
        sides = csr if csr is not None else [cs.csr(k) for k in range(3)]
        self.mats = [ctx.sparse_matrix(self.F.name, nc, cs.num_variables + 1, *sides[k]) for k in range(3)]


def witness_map(ctx, pk, x, check_satisfied=False):
    """r1cs_to_qap<F>::witness_map with d1 = d2 = d3 = 0 (prover.hpp:78-82): x = (1, primary, auxiliary) on the device
    -> coefficients_for_H [m, 8] (the entries m-1 and m of the reference's m+1 vector are zero by construction).
    check_satisfied: the reference's precondition `constraint_system.is_satisfied(primary, auxiliary)` (prover.hpp:77,
    r1cs_to_qap.hpp:225) on the row values <a_i, x> <b_i, x> - <c_i, x> the map computes anyway; raises when it fails
    (without it an unsatisfying assignment silently yields a garbage proof: the division by Z drops the remainder)."""
    import torch
    F, m, log_m = pk.F, pk.m, pk.log_m
    cs = pk.cs
    nc, ni = cs.num_constraints, cs.num_inputs
    if tuple(x.shape) != (cs.num_variables + 1, 8):
        raise capi.ZkbInvalidArgument(capi.ERR_INVALID_ARGUMENT, "assignment must be [num_variables + 1, 8] limbs (1, primary, auxiliary)")
    abc = torch.zeros((3, m, 8), dtype=torch.int32, device=x.device)
    for k in range(3):
        pk.mats[k].matvec(x, abc[k])
    if check_satisfied:
        one = 1
        resid = ctx.vec(F.name, capi.VEC_MUL_SUB_SCALE, abc[0, :nc], abc[1, :nc], abc[2, :nc], scalar=one)
        if bool(resid.any().item()) or not bool((x[0] == torch.tensor([1, 0, 0, 0, 0, 0, 0, 0], dtype=torch.int32, device=x.device)).all().item()):
            raise capi.ZkbInvalidArgument(capi.ERR_INVALID_ARGUMENT, "the assignment does not satisfy the constraint system (prover.hpp:77)")
    abc[0, nc:nc + ni + 1] = x[:ni + 1]          # the input-consistency constraints input_i * 0 = 0 (:239-242)
    ctx.ntt(F.name, abc, log_m, inverse=True)
    ctx.ntt(F.name, abc, log_m, coset_shift=F.generator)
    z_inv = pow((pow(F.generator, m, F.p) - 1) % F.p, F.p - 2, F.p)   # divide_by_z_on_coset: Z(g) = g^m - 1
    h = ctx.vec(F.name, capi.VEC_MUL_SUB_SCALE, abc[0], abc[1], abc[2], scalar=z_inv, out=abc[0])
    ctx.ntt(F.name, h.unsqueeze(0), log_m, inverse=True, coset_shift=F.generator)
    return h


def prove(ctx, pk, primary_input, auxiliary_input, r, s, x_device=None, concurrent=False, check_satisfied=True):
    """Returns (g1_A, g2_B, g1_C) in affine form.  r, s: the prover's zero-knowledge randomness (the reference draws
    them with algebra::random_element, prover.hpp:91-92).  x_device: optional device tensor with the full assignment
    (1, primary, auxiliary) as [num_variables + 1, 8] limbs; otherwise it is uploaded from the Python integers.
    concurrent: the five multiexps of prover.hpp:108-139 are independent; run each on its own stream (and context: a
    context's scratch belongs to one call at a time) from its own host thread, so the latency-bound bucket-reduce tail of
    one overlaps the bucket accumulation of the others.  Same proof."""
    import torch
    F = pk.F
    p = F.p
    cs = pk.cs
    nv, ni, m = cs.num_variables, cs.num_inputs, pk.m
    dev = "cuda:%d" % ctx.device
    if x_device is None:
        full = [1] + [int(v) % p for v in primary_input] + [int(v) % p for v in auxiliary_input]
        if len(full) != nv + 1:
            raise capi.ZkbInvalidArgument(capi.ERR_INVALID_ARGUMENT, "assignment size does not match the constraint system")
        x_device = torch.from_numpy(_int_rows(full).view(np.int32)).to(dev)
    x = x_device
    h = witness_map(ctx, pk, x, check_satisfied=check_satisfied)

    def head(*vals):
        return torch.from_numpy(_int_rows([v % p for v in vals]).view(np.int32)).to(dev)

    sa = torch.cat([head(1, r), x])
    xb = x.index_select(0, pk.B_index)
    sb = torch.cat([head(1, s), xb])
    sh, sl = h[:m - 1].contiguous(), x[ni + 1:].contiguous()
    jobs = [(pk.B2, sb, None), (pk.A, sa, None), (pk.B1, sb, None), (pk.H, sh, m - 1), (pk.L, sl, None)]
    if not concurrent:
        g2_B, g1_A, g1_B, ev_H, ev_L = [ctx.multiexp(b, sc, n=n) for b, sc, n in jobs]
    else:
        import threading
        from .api import Context
        if not hasattr(pk, "_side"):
            pk._side = [(Context(ctx.device), torch.cuda.Stream(device=dev)) for _ in jobs]
        main = torch.cuda.current_stream(dev)
        out, err = [None] * len(jobs), []

        def run(i):
            try:
                c, st = pk._side[i]
                torch.cuda.set_device(ctx.device)
                st.wait_stream(main)
                with torch.cuda.stream(st):
                    b, sc, n = jobs[i]
                    out[i] = c.multiexp(b, sc, n=n)
            except Exception as e:   # surfaces in the calling thread
                err.append(e)

        th = [threading.Thread(target=run, args=(i,)) for i in range(len(jobs))]
        for t_ in th:
            t_.start()
        for t_ in th:
            t_.join()
        if err:
            raise err[0]
        g2_B, g1_A, g1_B, ev_H, ev_L = out
    tail = ctx.msm_bases(pk.g1.name, _points_array(pk.g1, [ev_H, ev_L, g1_A, g1_B, pk.delta_g1]))
    g1_C = ctx.multiexp(tail, _int_rows([1, 1, s % p, r % p, (-r * s) % p]))
    tail.free()
    return g1_A, g2_B, g1_C


# ------------------------------------------------------------------------------------------------------------------
# r1cs_gg_ppzksnark_generator (generator.hpp:83-235): the QAP evaluated at t on the host, every group element by
# fixed-base batch exponentiation on the device (zkb_batch_exp = algebra::batch_exp / kc_batch_exp of :187-225).

def _batch_inverse(vals, p):
    """Montgomery's trick: one modular inversion for the whole list (no zero entries)."""
    pref, acc = [], 1
    for v in vals:
        pref.append(acc)
        acc = acc * v % p
    inv = pow(acc, p - 2, p)
    out = [0] * len(vals)
    for i in range(len(vals) - 1, -1, -1):
        out[i] = inv * pref[i] % p
        inv = inv * vals[i] % p
    return out


def _lagrange_at(F, log_m, t):
    """basic_radix2_domain::evaluate_all_lagrange_polynomials(t): L_i(t) over the size-2^log_m subgroup."""
    from .fields import omega as _omega
    p, m = F.p, 1 << log_m
    w = _omega(F, log_m)
    pows = [1] * m
    for i in range(1, m):
        pows[i] = pows[i - 1] * w % p
    if pow(t, m, p) == 1:                      # t is a domain element: the indicator vector
        return [1 if x == t % p else 0 for x in pows]
    l0 = (pow(t, m, p) - 1) * pow(m, p - 2, p) % p
    inv = _batch_inverse([(t - x) % p for x in pows], p)
    return [l0 * pows[i] % p * inv[i] % p for i in range(m)]


def qap_instance_evaluation(cs, F, t):
    """r1cs_to_qap::instance_map_with_evaluation (r1cs_to_qap.hpp:147-204): (At, Bt, Ct, Ht, Zt, m) with the
    input-consistency rows num_constraints + i added to A."""
    p = F.p
    m = cs.num_constraints + cs.num_inputs + 1
    if m & (m - 1):
        raise capi.ZkbInvalidArgument(capi.ERR_INVALID_ARGUMENT,
                                      "num_constraints + num_inputs + 1 must be a power of two (basic_radix2_domain)")
    u = _lagrange_at(F, m.bit_length() - 1, t)
    At, Bt, Ct = ([0] * (cs.num_variables + 1) for _ in range(3))
    for i in range(cs.num_inputs + 1):
        At[i] = u[cs.num_constraints + i]
    for row, con in enumerate(cs.constraints):
        for side, acc in zip(con, (At, Bt, Ct)):
            for idx, co in side:
                acc[idx] = (acc[idx] + u[row] * co) % p
    Ht = [1] * (m + 1)
    for i in range(1, m + 1):
        Ht[i] = Ht[i - 1] * t % p
    return At, Bt, Ct, Ht, (pow(t, m, p) - 1) % p, m


def swap_ab_if_beneficial(cs):
    """r1cs_constraint_system::swap_AB_if_beneficial (generator.hpp:88): B goes to G2, so the side with fewer distinct
    variables becomes B."""
    ta = {i for a, _, _ in cs.constraints for i, _ in a}
    tb = {i for _, b, _ in cs.constraints for i, _ in b}
    if len(tb) > len(ta):
        return R1csConstraintSystem(cs.num_inputs, cs.num_aux, [(b, a, c) for a, b, c in cs.constraints])
    return cs


def generator(ctx, curve_g1, curve_g2, cs, t, alpha, beta, gamma, delta, g1_generator=None, g2_generator=None):
    """basic_process with the toxic waste passed in (the reference draws t, alpha, beta, gamma, delta and the two
    generators at random, generator.hpp:92-102,152,160).  Returns (key, vk): `key` is a dict with the argument names of
    ProvingKey / marshalling.proving_key_to_bytes (points affine, None = infinity), `vk` holds gamma_g2, delta_g2,
    gamma_g1 and gamma_ABC_g1 = (first, rest); alpha_g1_beta_g2 is a pairing value and not computed on this path."""
    from .api import _affine_from_limbs
    g1 = CURVE_BY_NAME[curve_g1] if isinstance(curve_g1, str) else curve_g1
    g2 = CURVE_BY_NAME[curve_g2] if isinstance(curve_g2, str) else curve_g2
    F = FIELD_BY_NAME[g1.scalar_field]
    p = F.p
    t, alpha, beta, gamma, delta = (int(v) % p for v in (t, alpha, beta, gamma, delta))
    if gamma == 0 or delta == 0:
        raise capi.ZkbInvalidArgument(capi.ERR_INVALID_ARGUMENT, "gamma and delta must be invertible (generator.hpp:162-163)")
    cs = swap_ab_if_beneficial(cs)
    At, Bt, Ct, Ht, Zt, m = qap_instance_evaluation(cs, F, t)
    ginv, dinv = pow(gamma, p - 2, p), pow(delta, p - 2, p)
    ni, nv = cs.num_inputs, cs.num_variables
    abc = [(beta * At[i] + alpha * Bt[i] + Ct[i]) % p for i in range(nv + 1)]
    gamma_abc = [v * ginv % p for v in abc[:ni + 1]]
    Lt = [v * dinv % p for v in abc[ni + 1:]]
    coeff = Zt * dinv % p
    Hs = [coeff * h % p for h in Ht[:len(Ht) - 2]]
    b_idx = [i for i, v in enumerate(Bt) if v]
    b_sc = [Bt[i] for i in b_idx]

    def exp(curve, base, scalars):
        cl = coord_limbs(curve)
        base = (curve.gen_x, curve.gen_y) if base is None else base
        if not scalars:
            return []
        out = np.asarray(ctx.batch_exp(curve.name, base, _int_rows(scalars)), dtype=np.uint32)
        return [_affine_from_limbs(out[i].reshape(-1), cl, curve.deg) for i in range(len(scalars))]

    head1 = [alpha, beta, delta, gamma]
    pts1 = exp(g1, g1_generator, head1 + gamma_abc + At + b_sc + Hs + Lt)
    pts2 = exp(g2, g2_generator, [beta, delta, gamma] + b_sc)
    parts, o = [], len(head1)
    for n in (ni + 1, nv + 1, len(b_sc), len(Hs), len(Lt)):
        parts.append(pts1[o:o + n])
        o += n
    gabc_pts, A_query, B_g1, H_query, L_query = parts
    key = dict(alpha_g1=pts1[0], beta_g1=pts1[1], beta_g2=pts2[0], delta_g1=pts1[2], delta_g2=pts2[1],
               A_query=A_query, B_indices=b_idx, B_g2=pts2[3:], B_g1=B_g1, B_domain_size=nv + 1,
               H_query=H_query, L_query=L_query, num_inputs=ni, num_aux=cs.num_aux, constraints=cs.constraints)
    vk = dict(gamma_g2=pts2[2], delta_g2=pts2[1], gamma_g1=pts1[3], gamma_ABC_g1=(gabc_pts[0], gabc_pts[1:]))
    return key, vk


# ---------------------------------------------------------------------------------------------------------------------
# The generator with every vector on the device: nothing below touches a Python integer per element.
def csr_transpose(row_ptr, col_idx, values, ncols):
    """CSR arrays of the transposed matrix (numpy; stable in the row index, so every output row keeps constraint order)"""
    row_ptr = np.asarray(row_ptr, dtype=np.int64)
    col_idx = np.asarray(col_idx, dtype=np.int64)
    rows = np.repeat(np.arange(row_ptr.size - 1, dtype=np.int64), np.diff(row_ptr))
    order = np.argsort(col_idx, kind="stable")
    tp = np.zeros(ncols + 1, dtype=np.uint64)
    tp[1:] = np.cumsum(np.bincount(col_idx, minlength=ncols)).astype(np.uint64)
    return tp, rows[order].astype(np.uint32), np.ascontiguousarray(np.asarray(values)[order])


def _const_rows(ctx, value, n):
    import torch
    row = torch.from_numpy(_int_rows([int(value)]).view(np.int32)).to("cuda:%d" % ctx.device)
    return row.expand(n, 8).contiguous()


def qap_instance_evaluation_device(ctx, F, num_constraints, num_inputs, num_variables, csr_sides, t):
    """r1cs_to_qap::instance_map_with_evaluation (r1cs_to_qap.hpp:138-204) on the device: u = all Lagrange polynomials at
    t (powers of omega by a product scan, one batched inversion), At / Bt / Ct = transposed sparse mat-vec of the three
    sides with u (plus the input-consistency rows on A), Ht = powers of t.  csr_sides = [(row_ptr, col, values)] x 3 of
    the constraint system (rows = constraints).  Returns device tensors (At, Bt, Ct [num_variables + 1, 8],
    Ht [m + 1, 8]) and the integers Zt, m."""
    import torch
    from .fields import omega as _omega
    p = F.p
    nc, ni, nv = num_constraints, num_inputs, num_variables
    m = nc + ni + 1
    if m & (m - 1):
        raise capi.ZkbInvalidArgument(capi.ERR_INVALID_ARGUMENT,
                                      "num_constraints + num_inputs + 1 must be a power of two (basic_radix2_domain)")
    log_m = m.bit_length() - 1
    t = int(t) % p
    zt = (pow(t, m, p) - 1) % p
    if zt == 0:
        raise capi.ZkbInvalidArgument(capi.ERR_INVALID_ARGUMENT, "t lies in the evaluation domain (Z(t) = 0: no quotient, generator.hpp draws t at random)")
    pows = ctx.prefix_product(F.name, _const_rows(ctx, _omega(F, log_m), m))                 # omega^i
    inv = ctx.batch_inverse(F.name, ctx.vec(F.name, capi.VEC_SUB, _const_rows(ctx, t, m), pows))
    u = ctx.vec(F.name, capi.VEC_MUL, pows, inv)
    u = ctx.poly_lincomb(F.name, u, m, [zt * pow(m, p - 2, p) % p])                           # L_i(t) = Z(t)/m omega^i / (t - omega^i)
    sides = []
    x = u[:nc].contiguous()
    for row_ptr, col, vals in csr_sides:
        tp, tc, tv = csr_transpose(row_ptr, col, vals, nv + 1)
        mat = ctx.sparse_matrix(F.name, nv + 1, nc, tp, tc, tv)
        out = torch.empty((nv + 1, 8), dtype=torch.int32, device=u.device)
        mat.matvec(x, out)
        mat.free()
        sides.append(out)
    at = sides[0]
    at[:ni + 1] = ctx.vec(F.name, capi.VEC_ADD, at[:ni + 1].contiguous(), u[nc:nc + ni + 1].contiguous())
    ht = ctx.prefix_product(F.name, _const_rows(ctx, t, m + 1))
    return at, sides[1], sides[2], ht, zt, m


def generator_device(ctx, curve_g1, curve_g2, num_constraints, num_inputs, num_variables, csr_sides, t, alpha, beta, gamma,
                     delta, g1_generator=None, g2_generator=None):
    """r1cs_gg_ppzksnark_generator::basic_process (generator.hpp:83-235) with the QAP evaluation, the key scalars and every
    key element on the device; the caller applies swap_AB_if_beneficial to csr_sides beforehand (workloads does).
    Returns (key, vk) like `generator`, but the query vectors are device tensors [n, 2, limbs] ((0, 0) = infinity) and
    B_indices a numpy array - the form ProvingKey / proving_key_from_dict take."""
    import torch
    from .api import _affine_from_limbs
    g1 = CURVE_BY_NAME[curve_g1] if isinstance(curve_g1, str) else curve_g1
    g2 = CURVE_BY_NAME[curve_g2] if isinstance(curve_g2, str) else curve_g2
    F = FIELD_BY_NAME[g1.scalar_field]
    p = F.p
    t, alpha, beta, gamma, delta = (int(v) % p for v in (t, alpha, beta, gamma, delta))
    if gamma == 0 or delta == 0:
        raise capi.ZkbInvalidArgument(capi.ERR_INVALID_ARGUMENT, "gamma and delta must be invertible (generator.hpp:162-163)")
    ni, nv = num_inputs, num_variables
    at, bt, ct, ht, zt, m = qap_instance_evaluation_device(ctx, F, num_constraints, ni, nv, csr_sides, t)
    ginv, dinv = pow(gamma, p - 2, p), pow(delta, p - 2, p)
    abc = ctx.poly_lincomb(F.name, torch.stack([at, bt, ct]), nv + 1, [beta, alpha, 1])
    gamma_abc = ctx.poly_lincomb(F.name, abc[:ni + 1].contiguous(), ni + 1, [ginv])
    lt = ctx.poly_lincomb(F.name, abc[ni + 1:].contiguous(), nv - ni, [dinv]) if nv > ni else abc[:0]
    hs = ctx.poly_lincomb(F.name, ht[:m - 1].contiguous(), m - 1, [zt * dinv % p])
    b_idx = torch.nonzero((bt != 0).any(dim=1)).flatten()
    b_sc = bt.index_select(0, b_idx)
    head1 = torch.from_numpy(_int_rows([alpha, beta, delta, gamma]).view(np.int32)).to(at.device)
    head2 = torch.from_numpy(_int_rows([beta, delta, gamma]).view(np.int32)).to(at.device)
    base1 = (g1.gen_x, g1.gen_y) if g1_generator is None else g1_generator
    base2 = (g2.gen_x, g2.gen_y) if g2_generator is None else g2_generator
    pts1 = ctx.batch_exp(g1.name, base1, torch.cat([head1, gamma_abc, at, b_sc, hs, lt]).contiguous())
    pts2 = ctx.batch_exp(g2.name, base2, torch.cat([head2, b_sc]).contiguous())
    cl1, cl2 = coord_limbs(g1), coord_limbs(g2)
    h1 = pts1[:4].cpu().numpy().view(np.uint32)
    h2 = pts2[:3].cpu().numpy().view(np.uint32)
    pt1 = [_affine_from_limbs(h1[i].reshape(-1), cl1, g1.deg) for i in range(4)]
    pt2 = [_affine_from_limbs(h2[i].reshape(-1), cl2, g2.deg) for i in range(3)]
    parts, o = [], 4
    for n in (ni + 1, nv + 1, int(b_idx.numel()), m - 1, nv - ni):
        parts.append(pts1[o:o + n])
        o += n
    gabc, a_query, b_g1, h_query, l_query = parts
    key = dict(alpha_g1=pt1[0], beta_g1=pt1[1], beta_g2=pt2[0], delta_g1=pt1[2], delta_g2=pt2[1], A_query=a_query,
               B_indices=b_idx.cpu().numpy(), B_g2=pts2[3:], B_g1=b_g1, B_domain_size=nv + 1, H_query=h_query, L_query=l_query,
               num_inputs=ni, num_aux=nv - ni)
    vk = dict(gamma_g2=pt2[2], delta_g2=pt2[1], gamma_g1=pt1[3], gamma_ABC_g1=gabc)
    return key, vk


def proving_key_from_dict(ctx, curve_g1, curve_g2, key, precompute=False):
    """Device-resident ProvingKey from the dict `generator` returns / marshalling.proving_key_from_bytes reads."""
    cs = R1csConstraintSystem(key["num_inputs"], key["num_aux"], key["constraints"])
    return ProvingKey(ctx, curve_g1, curve_g2, cs, key["alpha_g1"], key["beta_g1"], key["beta_g2"], key["delta_g1"],
                      key["delta_g2"], key["A_query"], key["B_indices"], key["B_g2"], key["B_g1"], key["H_query"],
                      key["L_query"], precompute=precompute)
