"""Host mirror of lpc_commitment_scheme (zk/commitments/polynomial/lpc.hpp:66-200) and its base
polys_evaluator (zk/commitments/batched_commitment.hpp:60-250) over device-resident polynomials.

What runs where:
  commit(index)        zkb_lpc_commit: LDE of the batch to D[0] + leaf hash + Merkle tree        (lpc.hpp:101-106)
  eval_polys()         batched inverse NTT + zkb_poly_evaluate                                      (batched_commitment.hpp:176-190)
  proof_eval()         theta powers on the host; per unique point zkb_poly_lincomb + zkb_poly_div_linear;
                       from_coefficients + resize to D[0] = one zero-padded NTT; then zkb_fri_commit_phase
                       (lpc.hpp:113-200, basic_fri.hpp:706-737)
  grinding             zkb_pow_grind: every thread tries one nonce                                   (proof_of_work.hpp:47-68)
  query phase          challenges and index recurrences on the host; the opened values by one zkb_poly_evaluate
                       per batch over all query points, the round values by one gather from the retained fs,
                       every tree's lambda paths by one zkb_merkle_paths                            (basic_fri.hpp:749-915)
proof_eval returns the evaluation table z, the combined Q on D[0], the commit-phase outputs and - with query=True -
the complete lpc proof {z, fri_proof} in the data model of oracle/fri_query.py.  Proof marshalling is not built.
Polynomials are polynomial_dfs values: evaluations on the 2^k subgroup, canonical uint32 limbs, torch CUDA tensors.
"""
import numpy as np

from . import capi
from .api import _int_rows
from .fields import FIELD_BY_NAME, omega


class FriParams:
    """commitments::detail::basic_batched_fri::params_type (basic_fri.hpp:151-183): r = max degree log,
    D[i] of size 2^(r + expand_factor - i) (calculate_domain_set), step_list summing to r."""

    def __init__(self, step_list, degree_log, lambda_=40, expand_factor=2, use_grinding=False, grinding_parameter=0xFFFF):
        self.use_grinding = bool(use_grinding)
        self.grinding_parameter = int(grinding_parameter)
        self.step_list = [int(s) for s in step_list]
        self.r = sum(self.step_list)
        self.degree_log = int(degree_log)
        self.lambda_ = int(lambda_)
        self.expand_factor = int(expand_factor)
        self.log_d0 = self.degree_log + self.expand_factor
        if self.r > self.log_d0:
            raise ValueError("sum(step_list) exceeds log2 |D[0]|")

    @staticmethod
    def with_max_step_one(degree_log, lambda_=40, expand_factor=2, use_grinding=False, grinding_parameter=0xFFFF):
        """params_type(max_step = 1, degree_log, lambda, expand_factor) (basic_fri.hpp:150-166), the constructor the
        reference's Placeholder tests use (test/systems/plonk/placeholder/placeholder.cpp:231): r = degree_log - 1
        rounds of step 1 (generate_random_step_list is deterministic for max_step = 1)."""
        return FriParams([1] * (degree_log - 1), degree_log, lambda_, expand_factor, use_grinding, grinding_parameter)


class LpcCommitmentScheme:
    def __init__(self, ctx, field, hash_id, fri_params, retain_lde=False):
        """retain_lde: keep the extended evaluations of every committed batch on the device (|D0| elements per polynomial).
        The reference drops them and re-evaluates the polynomials at the 2 lambda query points instead, "it takes waaay too
        much RAM" (basic_fri.hpp:689-691, 750-775); with 180 GB of HBM the 2^20-row Placeholder circuit's 101 columns at
        blow-up 8 are 27 GB, and the query phase becomes a gather."""
        self.ctx, self.hash_id, self.fri = ctx, hash_id, fri_params
        self.retain_lde = bool(retain_lde)
        self._ext = {}
        self._commit_coeffs = {}
        self._cache = {}
        self._whole = {}
        self.F = FIELD_BY_NAME[field] if isinstance(field, str) else field
        self._polys = {}       # batch index -> list of [n, 8] device tensors (same n within a batch)
        self._points = {}      # batch index -> list (per polynomial) of point lists
        self._trees = {}
        self._fixed = {}
        self._locked = {}
        self._etha = None
        self._fixed_values = None
        self.z = {}
        self.timings = {}

    # ---- polys_evaluator interface (batched_commitment.hpp:196-250)
    def append_to_batch(self, index, poly):
        if self._locked.get(index):
            raise RuntimeError("batch %d is already committed" % index)
        t = poly if poly.dim() == 3 else poly.unsqueeze(0)
        # a batch handed over as one [count, n, 8] tensor is used as it is (no re-stacking copy per commit / eval_polys)
        self._whole[index] = t.contiguous() if index not in self._polys else None
        for i in range(t.shape[0]):
            self._polys.setdefault(index, []).append(t[i])

    def append_eval_point(self, batch, point, poly=None):
        """(batch, point): every polynomial of the batch; (batch, poly, point): one polynomial."""
        if batch not in self._points:      # a prover-side batch; a verifier calls set_batch_size first
            self._points[batch] = [[] for _ in self._polys[batch]]
        pts = self._points[batch]
        targets = range(len(pts)) if poly is None else [poly]
        for i in targets:
            pts[i].append(int(point) % self.F.p)

    def set_batch_size(self, index, size):
        self._points[index] = [[] for _ in range(size)]

    def _batch_tensor(self, index):
        import torch
        polys = self._polys[index]
        n = polys[0].shape[0]
        if any(p.shape[0] != n for p in polys):
            raise ValueError("polynomials of one batch must have the same size")
        if self._whole.get(index) is None:
            self._whole[index] = torch.stack(polys).contiguous()
        return self._whole[index], n

    # ---- lpc_commitment_scheme (lpc.hpp:95-111)
    def commit(self, index):
        self._locked[index] = True
        self._points.setdefault(index, [[] for _ in self._polys[index]])
        batch, n = self._batch_tensor(index)
        log_n = n.bit_length() - 1
        if self.retain_lde:
            import torch
            co = torch.empty_like(batch)      # the coefficient form is a by-product of the resize: eval_polys reuses it
            ext = self.ctx.lde(self.F.name, batch, log_n, self.fri.log_d0, coefficients_out=co)
            self._commit_coeffs[index] = (co, n)
            tree = self.ctx.merkle_commit(self.F.name, self.hash_id, ext, self.fri.log_d0, self.fri.step_list[0], keep_tree=True)
            self._ext[index] = ext
        else:
            tree = self.ctx.lpc_commit(self.F.name, self.hash_id, batch, log_n, self.fri.log_d0, self.fri.step_list[0],
                                       keep_tree=True)
        self._trees[index] = tree
        return tree.root()

    def mark_batch_as_fixed(self, index):
        self._fixed[index] = True

    def has_batch(self, index):
        return index in self._polys

    def fixed_batch_values(self, index):
        """values of a fixed batch's polynomials at etha, evaluated the way eval_polys evaluates (the reference's
        preprocessor stores them in the commitment scheme's preprocessed data, lpc.hpp:83-93); call after setup()"""
        batch, n = self._batch_tensor(index)
        co = self.ctx.ntt(self.F.name, batch.clone(), n.bit_length() - 1, inverse=True)
        vals = [v[0] for v in self.ctx.poly_evaluate(self.F.name, co, n, [self._etha])]
        if self._fixed_values is None:
            self._fixed_values = {}
        self._fixed_values[index] = vals
        return vals

    def setup(self, transcript, fixed_values):
        self._etha = transcript.challenge(self.F.p)
        self._fixed_values = fixed_values

    # ---- eval_polys (batched_commitment.hpp:176-190)
    def eval_polys(self):
        self._coeffs = {}
        self.z = {}
        for k in sorted(self._polys):
            if k in self._commit_coeffs:
                co, n = self._commit_coeffs[k]
            else:
                batch, n = self._batch_tensor(k)
                co = self.ctx.workspace("lpc_coeffs_%d" % k, batch.shape)     # valid until the next proof on this context
                self.ctx.ntt(self.F.name, batch, n.bit_length() - 1, inverse=True, out=co)
            self._coeffs[k] = (co, n)
            union = []
            for pts in self._points[k]:
                for x in pts:
                    if x not in union:
                        union.append(x)
            if not union:
                self.z[k] = [[] for _ in self._points[k]]
                continue
            vals = self.ctx.poly_evaluate(self.F.name, co, n, union)
            self.z[k] = [[vals[i][union.index(x)] for x in pts] for i, pts in enumerate(self._points[k])]
        return self.z

    def _unique_points(self):
        out = []
        for k in sorted(self._points):
            for pts in self._points[k]:
                for x in pts:
                    if x not in out:
                        out.append(x)
        return out

    # ---- proof_eval up to the end of the FRI commit phase (lpc.hpp:113-200)
    def proof_eval(self, transcript, keep_trees=False, keep_fs=False, query=False):
        import torch
        if query:
            keep_trees = keep_fs = True
        p = self.F.p
        self.eval_polys()
        for k in sorted(self._trees):
            transcript(self._trees[k].root())
        theta = transcript.challenge(p)
        theta_acc = 1
        n_max = max(n for _, n in self._coeffs.values())
        dev = next(iter(self._coeffs.values()))[0].device
        combined = self.ctx.workspace("lpc_combined", (n_max, 8)).zero_()
        numer = self.ctx.workspace("lpc_numer", (n_max, 8))
        quot = self.ctx.workspace("lpc_quot", (n_max, 8))
        remainders = []

        def add_quotient(point, terms):
            """terms: list of (batch index, scalars per polynomial, constant)"""
            numer.zero_()
            for k, scalars, constant in terms:
                co, n = self._coeffs[k]
                self.ctx.poly_lincomb(self.F.name, co, n, scalars, constant=constant, out=numer, accumulate=True)
            _, rem = self.ctx.poly_div_linear(self.F.name, numer, n_max, point, out=quot)
            remainders.append(rem)
            self.ctx.vec(self.F.name, capi.VEC_ADD, combined, quot, out=combined)

        for point in self._unique_points():
            terms = []
            for k in sorted(self._polys):
                scalars, constant, used = [], 0, False
                for i, pts in enumerate(self._points[k]):
                    if point not in pts:
                        scalars.append(0)
                        continue
                    j = pts.index(point)
                    scalars.append(theta_acc)
                    constant = (constant + self.z[k][i][j] * theta_acc) % p
                    theta_acc = theta_acc * theta % p
                    used = True
                if used:
                    terms.append((k, scalars, constant))
            add_quotient(point, terms)
        for k in sorted(self._polys):
            if not self._fixed.get(k):
                continue
            scalars, constant = [], 0
            for i in range(len(self._polys[k])):
                scalars.append(theta_acc)
                constant = (constant + self._fixed_values[k][i] * theta_acc) % p
                theta_acc = theta_acc * theta % p
            add_quotient(self._etha, [(k, scalars, constant)])
        # combined_Q.from_coefficients(combined_Q_normal) and precommit's resize to D[0]: the same polynomial
        # evaluated on D[0] - one forward NTT of the zero-padded coefficients
        nd = 1 << self.fri.log_d0
        q_d0 = self.ctx.workspace("lpc_q_d0", (1, nd, 8)).zero_()
        q_d0[0, :n_max] = combined
        self.ctx.ntt(self.F.name, q_d0, self.fri.log_d0)
        fri = self.ctx.fri_commit_phase(
            self.F.name, self.hash_id, q_d0[0], self.fri.log_d0, self.fri.step_list,
            lambda rnd, root, count: (transcript(root), [transcript.challenge(p) for _ in range(count)])[1],
            keep_trees=keep_trees, keep_fs=keep_fs)
        out = {"z": self.z, "theta": theta, "combined_Q_normal": combined, "combined_Q": q_d0[0],
               "remainders": remainders, "fri": fri}
        if query:
            import time
            t0 = time.perf_counter()
            out["proof"] = {"z": self.z, "fri_proof": self._query_phase(transcript, fri)}
            self.timings["query_phase_ms"] = (time.perf_counter() - t0) * 1e3
        return out

    # ---- verify_eval (lpc.hpp:202-263): host code, like the reference's; a verifier-side scheme only needs
    # set_batch_size / append_eval_point (and setup for fixed batches) before this call
    def verify_eval(self, proof, commitments, transcript):
        from . import lpc_verify
        fixed = tuple(k for k, v in self._fixed.items() if v)
        return lpc_verify.lpc_verify_eval(self.F, self.hash_id, self.fri, proof, self._points, commitments, transcript,
                                          fixed, self._etha, self._fixed_values)

    # ---- grinding + query phase of zk::algorithms::proof_eval<FRI> (basic_fri.hpp:743-915)
    def _domain_index(self, x, log_n):
        """index of x in the 2^log_n subgroup (the reference searches linearly, basic_fri.hpp:780-786): bit by bit,
        bit i of the exponent is set iff (x w^-e)^(2^(log_n-1-i)) != 1 for the bits e found so far"""
        p = self.F.p
        key = ("winv", log_n)
        if key not in self._cache:
            w_inv, tab = pow(omega(self.F, log_n), p - 2, p), []
            for _ in range(log_n):
                tab.append(w_inv)              # w^-(2^i)
                w_inv = w_inv * w_inv % p
            self._cache[key] = tab
        tab = self._cache[key]
        e, cur = 0, x % p
        for i in range(log_n):
            if pow(cur, 1 << (log_n - 1 - i), p) != 1:
                e |= 1 << i
                cur = cur * tab[i] % p
        return e

    @staticmethod
    def _s_indices(x_index, domain_size, fri_step):
        """index pairs of calculate_s (basic_fri.hpp:583-617)"""
        half = domain_size // 2
        idx = [x_index]
        base, prev = domain_size // 4, 1
        while len(idx) < (1 << fri_step) // 2:
            idx += [(base + idx[j]) % domain_size for j in range(prev)]
            base //= 2
            prev <<= 1
        return [(a, (a + half) % domain_size) for a in idx]

    @staticmethod
    def _folded_index(x_index, domain_size, fri_step):
        for _ in range(fri_step):
            domain_size //= 2
            x_index %= domain_size
        return x_index

    def _query_phase(self, transcript, fri):
        import torch
        F, p, steps, log_d0 = self.F, self.F.p, self.fri.step_list, self.fri.log_d0
        if steps[-1] != 1:
            raise ValueError("step_list must end with 1 (check_step_list, basic_fri.hpp:544-571)")
        proof = {"fri_roots": fri["roots"], "final_polynomial": fri["final_polynomial"], "proof_of_work": None}
        if self.fri.use_grinding:
            nonce = self.ctx.pow_grind(transcript.hash_id, transcript.state, self.fri.grinding_parameter)
            transcript(nonce.to_bytes(4, "big"))
            transcript.int_challenge(32)
            proof["proof_of_work"] = nonce
        lam, d0 = self.fri.lambda_, 1 << log_d0
        # the challenges do not depend on anything opened, so draw them all and batch the device work
        x_idx0 = []
        for _ in range(lam):
            x = pow(transcript.challenge(p), (p - 1) // d0, p)
            x_idx0.append(self._domain_index(x, log_d0))
        w0 = omega(F, log_d0)
        init_pairs = [[(min(a, b), max(a, b)) for a, b in self._s_indices(xi, d0, steps[0])] for xi in x_idx0]
        # every pair is (i, i + |D0|/2): the points z = w^i and -z, which zkb_poly_evaluate_pm opens in one pass
        half0 = d0 // 2
        flat = sorted({a for pairs in init_pairs for a, _ in pairs})
        pos = {i: n for n, i in enumerate(flat)}
        pts = [pow(w0, i, p) for i in flat]
        initial = [dict() for _ in range(lam)]
        leaf0 = [self._folded_index(xi, d0, steps[0]) for xi in x_idx0]
        for k in sorted(self._polys):
            co, n = self._coeffs[k]
            if n == d0 or k in self._ext:    # on D[0] already (basic_fri.hpp:812-818) or retained: the values themselves
                batch = self._ext[k] if k in self._ext else self._batch_tensor(k)[0]
                idx = [i for a in flat for i in (a, a + half0)]
                sel = batch[:, torch.tensor(idx, device=batch.device)].cpu().numpy().view(np.uint32)
                vals = [[(int.from_bytes(sel[i, 2 * j].tobytes(), "little"), int.from_bytes(sel[i, 2 * j + 1].tobytes(), "little"))
                         for j in range(len(flat))] for i in range(sel.shape[0])]
            else:          # evaluate the coefficient form at the 2 lambda points (:819-834)
                vals = self.ctx.poly_evaluate_pm(F.name, co, n, pts)
            paths = self._trees[k].paths(leaf0)
            root = self._trees[k].root()
            for q in range(lam):
                initial[q][k] = {"values": [[list(v[pos[a]]) for a, _ in init_pairs[q]] for v in vals],
                                 "p": {"index": leaf0[q], "path": paths[q], "root": root}}
        # round proofs: paths from every fri tree, y from the retained fs (one gather), final round from final_polynomial
        fs, fs_off, acc, off = fri["fs"], [], log_d0, 0
        for s_ in steps:
            acc -= s_
            fs_off.append((off, acc))
            off += 1 << acc
        rounds = [[None] * len(steps) for _ in range(lam)]
        gather = []
        t = 0
        xs = list(x_idx0)
        for i, s_ in enumerate(steps):
            size_t = 1 << (log_d0 - t)
            xs = [xi % size_t for xi in xs]
            leaves = [self._folded_index(xi, size_t, s_) for xi in xs]
            paths = fri["trees"][i].paths(leaves)
            t += s_
            size_n = 1 << (log_d0 - t)
            for q in range(lam):
                rp = {"p": {"index": leaves[q], "path": paths[q], "root": fri["roots"][i]}}
                if i < len(steps) - 1:
                    xq = xs[q] % size_n
                    pairs = [(min(a, b), max(a, b)) for a, b in self._s_indices(xq, size_n, steps[i + 1])]
                    rp["y"] = [[None, None] for _ in pairs]
                    for j, (a, b) in enumerate(pairs):
                        gather.append((fs_off[i][0] + a, rp["y"][j], 0))
                        gather.append((fs_off[i][0] + b, rp["y"][j], 1))
                else:
                    size_p = 1 << (log_d0 - t + 1)               # D[t-1]
                    xq = xs[q] % size_p
                    x = pow(omega(F, log_d0 - t + 1), xq, p)
                    x = x * x % p
                    ind = 0 if xq % (size_p // 2) < size_p // 4 else 1
                    y = [[0, 0]]
                    y[0][ind] = _horner(fri["final_polynomial"], x, p)
                    y[0][1 - ind] = _horner(fri["final_polynomial"], (p - x) % p, p)
                    rp["y"] = y
                rounds[q][i] = rp
            if i < len(steps) - 1:
                xs = [xi % size_n for xi in xs]
        if gather:
            sel = fs[torch.tensor([g[0] for g in gather], device=fs.device)].cpu().numpy().view(np.uint32)
            for n_, (_, tgt, slot) in enumerate(gather):
                tgt[slot] = int.from_bytes(sel[n_].tobytes(), "little")
        proof["query_proofs"] = [{"initial_proof": initial[q], "round_proofs": rounds[q]} for q in range(lam)]
        return proof


def _horner(coeffs, x, p):
    acc = 0
    for c in reversed(coeffs):
        acc = (acc * x + c) % p
    return acc


def host_poly(vals):
    """list of integers -> [n, 8] uint32 array (helper for callers that hold Python integers)"""
    return np.ascontiguousarray(_int_rows(vals))
