"""Python host layer over the C ABI: the names mirror the reference's call-site API
(math::evaluation_domain fft/inverse_fft, math::multiply_by_coset, polynomial_dfs::resize,
commitments::detail::fold_polynomial, zk::algorithms::precommit / lpc commit, algebra::multiexp).

Buffers are either numpy uint32 arrays (host; copied by the library inside the call) or torch CUDA
tensors (device resident; int32/uint32, last dimension = limbs).  Field elements are canonical
little-endian uint32 limbs.  PyTorch is used only as the owner of device memory and streams.
"""
import ctypes

import numpy as np

from . import capi
from .fields import CURVE_BY_NAME, FIELD_BY_NAME, coord_limbs


def _field_id(field):
    return FIELD_BY_NAME[field].fid if isinstance(field, str) else int(field)


def _curve(curve):
    return CURVE_BY_NAME[curve] if isinstance(curve, str) else curve


def _is_torch(x):
    return type(x).__module__.startswith("torch")


def _limbs(val, n):
    return (ctypes.c_uint32 * n)(*[(int(val) >> (32 * i)) & 0xFFFFFFFF for i in range(n)])


def _int_rows(vals, n=8):
    """list of integers -> [len, n] uint32 limb array"""
    blob = b"".join(int(v).to_bytes(4 * n, "little") for v in vals)
    return np.frombuffer(blob, dtype=np.uint32).reshape(len(vals), n).copy()


def _ints(arr):
    """[len, n] uint32 limb array -> list of integers"""
    a = np.ascontiguousarray(np.asarray(arr, dtype=np.uint32).reshape(-1, arr.shape[-1]))
    nb = 4 * a.shape[1]
    raw = a.tobytes()
    return [int.from_bytes(raw[i * nb:(i + 1) * nb], "little") for i in range(a.shape[0])]


class _Buf:
    """Pointer + location of a caller buffer."""

    def __init__(self, x, writable=False):
        if _is_torch(x):
            if not x.is_cuda:
                raise ValueError("torch tensors must live on the GPU (pass numpy arrays for host data)")
            if not x.is_contiguous():
                raise ValueError("tensor must be contiguous")
            if x.element_size() != 4:
                raise ValueError("tensor must have a 32-bit integer dtype")
            self.ptr, self.mem, self.nbytes = x.data_ptr(), capi.MEM_DEVICE, x.numel() * 4
        else:
            if x.dtype != np.uint32 or not x.flags["C_CONTIGUOUS"]:
                raise ValueError("host buffers must be C-contiguous numpy uint32 arrays")
            if writable and not x.flags["WRITEABLE"]:
                raise ValueError("output array is read-only")
            self.ptr, self.mem, self.nbytes = x.ctypes.data, capi.MEM_HOST, x.nbytes
        self.obj = x


def _empty_like(x, shape):
    if _is_torch(x):
        import torch
        return torch.empty(shape, dtype=x.dtype, device=x.device)
    return np.empty(shape, dtype=np.uint32)


def _stream_ptr(x, stream):
    if stream is not None:
        return ctypes.c_void_p(int(stream))
    if _is_torch(x):
        import torch
        return ctypes.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
    return ctypes.c_void_p(0)


class MerkleTree:
    """containers::merkle_tree<Hash,2> kept on the device."""

    def __init__(self, ctx, handle, root, digest_bytes):
        self._ctx, self._h, self._root, self.digest_bytes = ctx, handle, root, digest_bytes

    def root(self):
        return self._root

    @property
    def leaves(self):
        return int(capi.lib().zkb_merkle_leaves(self._h))

    def path(self, index):
        depth = self.leaves.bit_length() - 1
        buf = (ctypes.c_uint8 * (max(depth, 1) * self.digest_bytes))()
        capi.check(capi.lib().zkb_merkle_path(self._ctx._h, self._h, index, buf), self._ctx._h)
        raw = bytes(buf)
        return [raw[i * self.digest_bytes:(i + 1) * self.digest_bytes] for i in range(depth)]

    def paths(self, indices):
        """Authentication paths of several leaves with one device gather (FRI query phase)."""
        depth = self.leaves.bit_length() - 1
        count = len(indices)
        if count == 0 or depth == 0:
            return [[] for _ in range(count)]
        idx = (ctypes.c_uint64 * count)(*[int(i) for i in indices])
        buf = (ctypes.c_uint8 * (count * depth * self.digest_bytes))()
        capi.check(capi.lib().zkb_merkle_paths(self._ctx._h, self._h, count, idx, buf, None), self._ctx._h)
        raw, db = bytes(buf), self.digest_bytes
        return [[raw[(q * depth + d) * db:(q * depth + d + 1) * db] for d in range(depth)] for q in range(count)]

    def free(self):
        if self._h:
            capi.lib().zkb_merkle_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class SparseMatrix:
    """One side (A, B or C) of an r1cs_constraint_system as a device-resident CSR matrix."""

    def __init__(self, ctx, field, rows, cols, row_ptr, col_idx, values, stream=None):
        self._ctx, self.field, self.rows, self.cols = ctx, field, int(rows), int(cols)
        rp = np.ascontiguousarray(row_ptr, dtype=np.uint64)
        ci = np.ascontiguousarray(col_idx, dtype=np.uint32)
        va = np.ascontiguousarray(values, dtype=np.uint32)
        if rp.shape != (self.rows + 1,) or va.size != ci.size * 8:
            raise ValueError("CSR arrays: row_ptr [rows+1], col_idx [nnz], values [nnz, 8]")
        self._h = ctypes.c_void_p()
        capi.check(capi.lib().zkb_sparse_matrix_create(
            ctx._h, _field_id(field), self.rows, self.cols, rp.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)),
            capi.u32_ptr(ci) if ci.size else None, capi.u32_ptr(va) if va.size else None,
            ctypes.c_void_p(int(stream)) if stream is not None else None, ctypes.byref(self._h)), ctx._h)

    def matvec(self, x, out, stream=None):
        """out[i] = <row i, x>; x: [cols, 8] host array or device tensor, out: device tensor with >= rows elements."""
        bx, bo = _Buf(x), _Buf(out, writable=True)
        if bx.nbytes != self.cols * 32 or bo.mem != capi.MEM_DEVICE or bo.nbytes < self.rows * 32:
            raise ValueError("x must hold `cols` elements and out (device) at least `rows`")
        capi.check(capi.lib().zkb_sparse_matvec(self._ctx._h, self._h, bx.ptr, bx.mem, bo.ptr, _stream_ptr(out, stream)),
                   self._ctx._h)
        return out

    def free(self):
        if self._h:
            capi.lib().zkb_sparse_matrix_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class MsmBases:
    """A long-lived base vector (KZG commitment key, Groth16 A/B/H/L query) resident on the GPU."""

    def __init__(self, ctx, curve, points, stream=None):
        self._ctx, self.curve = ctx, _curve(curve)
        cl = coord_limbs(self.curve)
        b = _Buf(points)
        n = b.nbytes // (2 * cl * 4)
        if n * 2 * cl * 4 != b.nbytes:
            raise ValueError("points must be [n, 2, %d] uint32" % cl)
        h = ctypes.c_void_p()
        capi.check(capi.lib().zkb_msm_bases_create(ctx._h, self.curve.cid, n, b.ptr, b.mem, _stream_ptr(points, stream),
                                                   ctypes.byref(h)), ctx._h)
        self._h, self.n, self.coord_limbs, self.deg = h, n, cl, self.curve.deg

    def precompute(self, window_bits=0, max_bytes=8 << 30, stream=None):
        """Builds the window table 2^(c w) P_i once (zkb_msm_bases_precompute): later multiexps over these
        bases use ceil((bits+1)/c) windows that share one bucket set.  Returns self."""
        capi.check(capi.lib().zkb_msm_bases_precompute(self._ctx._h, self._h, window_bits, max_bytes,
                                                       _stream_ptr(None, stream)), self._ctx._h)
        return self

    def window_plan(self, n=None):
        """(widest window in bits, digit windows, bucket sets) zkb_msm uses for n scalars on these bases"""
        c, w, b = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        capi.check(capi.lib().zkb_msm_window_plan(self._h, self.n if n is None else n, ctypes.byref(c), ctypes.byref(w),
                                                  ctypes.byref(b)), self._ctx._h)
        return c.value, w.value, b.value

    def free(self):
        if self._h:
            capi.lib().zkb_msm_bases_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Context:
    """Per-GPU context (twiddle caches, scratch).  Raises if no CUDA device: there is no CPU path."""

    def __init__(self, device=0):
        self._h = ctypes.c_void_p()
        capi.check(capi.lib().zkb_ctx_create(device, ctypes.byref(self._h)))
        self.device = device

    def close(self):
        if self._h:
            capi.lib().zkb_ctx_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def workspace(self, name, shape):
        """A cached int32 device tensor of this shape, owned by the context and reused by later calls under the same name:
        a prover that builds one proof after another should not go back to the allocator for its multi-GB temporaries
        (a cudaMalloc inside a commit costs tens of milliseconds now and then).  Contents are undefined."""
        import torch
        if not hasattr(self, "_ws"):
            self._ws = {}
        shape = tuple(int(v) for v in shape)
        t = self._ws.get(name)
        if t is None or tuple(t.shape) != shape:
            self._ws[name] = None
            t = torch.empty(shape, dtype=torch.int32, device="cuda:%d" % self.device)
            self._ws[name] = t
        return t

    def kernel_launches(self):
        return int(capi.lib().zkb_ctx_kernel_launches(self._h))

    def set_scratch_limit(self, nbytes):
        capi.check(capi.lib().zkb_ctx_set_scratch_limit(self._h, nbytes), self._h)

    def release_caches(self):
        capi.check(capi.lib().zkb_ctx_release_caches(self._h), self._h)

    # ------------------------------------------------------------------ NTT / LDE
    def ntt(self, field, data, log_n, inverse=False, coset_shift=None, out=None, stream=None):
        """evaluation_domain<F>(2^log_n)::fft / inverse_fft over data[batch, 2^log_n, 8]; with
        coset_shift g: multiply_by_coset(a, g); fft(a)  /  inverse_fft(a); multiply_by_coset(a, g^-1)."""
        fid = _field_id(field)
        n = 1 << log_n
        b = _Buf(data)
        batch = b.nbytes // (n * 32)
        if batch * n * 32 != b.nbytes:
            raise capi.ZkbInvalidArgument(capi.ERR_INVALID_ARGUMENT, "data is not [batch, 2^log_n, 8] uint32")
        out = data if out is None else out
        o = _Buf(out, writable=True)
        if o.mem != b.mem or o.nbytes != b.nbytes:
            raise ValueError("out must match data in location and size")
        sh = _limbs(coset_shift, 8) if coset_shift is not None else None
        capi.check(capi.lib().zkb_ntt(self._h, fid, log_n, batch, b.ptr, o.ptr, int(bool(inverse)), sh, b.mem,
                                      _stream_ptr(data, stream)), self._h)
        return out

    def lde(self, field, data, log_n_in, log_n_out, out=None, stream=None, coefficients_out=None):
        """polynomial_dfs::resize(2^log_n_out) for data[batch, 2^log_n_in, 8] -> [batch, 2^log_n_out, 8].
        coefficients_out (device tensor like data): also receives the coefficient form (zkb_lde_with_coefficients)."""
        fid = _field_id(field)
        b = _Buf(data)
        batch = b.nbytes // ((1 << log_n_in) * 32)
        if batch * (1 << log_n_in) * 32 != b.nbytes:
            raise capi.ZkbInvalidArgument(capi.ERR_INVALID_ARGUMENT, "data is not [batch, 2^log_n_in, 8] uint32")
        if out is None:
            out = _empty_like(data, (batch, 1 << log_n_out, 8))
        o = _Buf(out, writable=True)
        if o.mem != b.mem or o.nbytes != batch * (1 << log_n_out) * 32:
            raise ValueError("out must be [batch, 2^log_n_out, 8] in the same memory space as data")
        if coefficients_out is not None:
            c = _Buf(coefficients_out, writable=True)
            if b.mem != capi.MEM_DEVICE or c.mem != capi.MEM_DEVICE or c.nbytes != b.nbytes:
                raise ValueError("coefficients_out must be a device tensor of the shape of data")
            capi.check(capi.lib().zkb_lde_with_coefficients(self._h, fid, log_n_in, log_n_out, batch, b.ptr, o.ptr, c.ptr,
                                                            _stream_ptr(data, stream)), self._h)
            return out
        capi.check(capi.lib().zkb_lde(self._h, fid, log_n_in, log_n_out, batch, b.ptr, o.ptr, b.mem,
                                      _stream_ptr(data, stream)), self._h)
        return out

    def vec(self, field, op, a, b, c=None, scalar=None, out=None, stream=None):
        fid = _field_id(field)
        ba, bb = _Buf(a), _Buf(b)
        n = ba.nbytes // 32
        out = _empty_like(a, tuple(a.shape)) if out is None else out
        o = _Buf(out, writable=True)
        cp = _Buf(c).ptr if c is not None else None
        sc = _limbs(scalar, 8) if scalar is not None else None
        capi.check(capi.lib().zkb_vec(self._h, fid, op, n, ba.ptr, bb.ptr, cp, sc, o.ptr, ba.mem, _stream_ptr(a, stream)),
                   self._h)
        return out

    # ------------------------------------------------------------------ FRI / LPC
    def fri_fold(self, field, f, log_n, alpha, out=None, stream=None):
        """commitments::detail::fold_polynomial (dfs form) on the domain of size 2^log_n."""
        fid = _field_id(field)
        b = _Buf(f)
        if b.nbytes != (1 << log_n) * 32:
            raise capi.ZkbInvalidArgument(capi.ERR_INVALID_ARGUMENT, "f is not [2^log_n, 8] uint32")
        if out is None:
            out = _empty_like(f, (1 << (log_n - 1), 8))
        o = _Buf(out, writable=True)
        capi.check(capi.lib().zkb_fri_fold(self._h, fid, log_n, b.ptr, _limbs(alpha, 8), o.ptr, b.mem,
                                           _stream_ptr(f, stream)), self._h)
        return out

    def lpc_commit(self, field, hash_id, polys, log_n_in, log_n_out, fri_step, keep_tree=False, stream=None):
        """zk::algorithms::precommit<FRI>(polys, D, fri_step) + root (lpc_commitment_scheme::commit).
        Returns root bytes, or a MerkleTree when keep_tree."""
        fid = _field_id(field)
        b = _Buf(polys)
        batch = b.nbytes // ((1 << log_n_in) * 32)
        db = capi.lib().zkb_merkle_digest_bytes(hash_id)
        root = (ctypes.c_uint8 * max(db, 1))()
        th = ctypes.c_void_p()
        capi.check(capi.lib().zkb_lpc_commit(self._h, fid, hash_id, log_n_in, log_n_out, fri_step, batch, b.ptr, b.mem,
                                             root, ctypes.byref(th) if keep_tree else None, _stream_ptr(polys, stream)),
                   self._h)
        return MerkleTree(self, th, bytes(root), db) if keep_tree else bytes(root)

    def merkle_commit(self, field, hash_id, evals, log_n, fri_step, keep_tree=False, stream=None):
        fid = _field_id(field)
        b = _Buf(evals)
        batch = b.nbytes // ((1 << log_n) * 32)
        db = capi.lib().zkb_merkle_digest_bytes(hash_id)
        root = (ctypes.c_uint8 * max(db, 1))()
        th = ctypes.c_void_p()
        capi.check(capi.lib().zkb_merkle_commit(self._h, fid, hash_id, log_n, fri_step, batch, b.ptr, b.mem, root,
                                                ctypes.byref(th) if keep_tree else None, _stream_ptr(evals, stream)),
                   self._h)
        return MerkleTree(self, th, bytes(root), db) if keep_tree else bytes(root)

    def merkle_root_of_digests(self, hash_id, digests):
        """Root of the binary tree over the given child digests (bytes, count a power of two): the top levels above
        per-GPU subtrees."""
        db = capi.lib().zkb_merkle_digest_bytes(hash_id)
        blob = b"".join(bytes(d) for d in digests)
        if len(blob) != db * len(digests):
            raise ValueError("digests must be %d bytes each" % db)
        root = (ctypes.c_uint8 * max(db, 1))()
        capi.check(capi.lib().zkb_merkle_root_of_digests(self._h, hash_id, len(digests), blob, root, None), self._h)
        return bytes(root)

    def fri_commit_phase(self, field, hash_id, f, log_n, step_list, challenge, keep_trees=False, keep_fs=False,
                         stream=None):
        """Commit phase of zk::algorithms::proof_eval<FRI> (basic_fri.hpp:706-737) on the device.
        f: combined Q in evaluation form on D[0] (2^log_n elements, device tensor or host array).
        challenge(round, root_bytes, count) -> list of `count` integers < p: the caller's transcript
        (transcript(root); alphas = transcript.challenge() x count).
        Returns a dict: roots, alphas (ints), final_polynomial (ints, coefficients), trees (MerkleTree per round or
        None), fs (device tensor with f after every round, back to back, or None)."""
        fid = _field_id(field)
        b = _Buf(f)
        if b.nbytes != (1 << log_n) * 32:
            raise capi.ZkbInvalidArgument(capi.ERR_INVALID_ARGUMENT, "f is not [2^log_n, 8] uint32")
        steps = [int(s) for s in step_list]
        rounds, total = len(steps), sum(steps)
        if rounds == 0 or min(steps) < 1 or total > log_n:
            raise capi.ZkbInvalidArgument(capi.ERR_INVALID_ARGUMENT, "step_list must be non-empty, >= 1 each, sum <= log_n")
        db = capi.lib().zkb_merkle_digest_bytes(hash_id)
        failure = []

        def _cb(user, rnd, root, root_bytes, count, out):
            try:
                vals = challenge(int(rnd), bytes(root[:root_bytes]), int(count))
                if len(vals) != count:
                    raise ValueError("challenge callback returned %d values, expected %d" % (len(vals), count))
                for k, v in enumerate(vals):
                    for l in range(8):
                        out[8 * k + l] = (int(v) >> (32 * l)) & 0xFFFFFFFF
                return 0
            except Exception as e:   # never unwind through the C frames
                failure.append(e)
                return 1

        cb = capi.FRI_CHALLENGE_FN(_cb)
        roots = (ctypes.c_uint8 * (db * rounds))()
        trees = (ctypes.c_void_p * rounds)() if keep_trees else None
        fs = None
        if keep_fs:
            import torch
            acc, n_fs = log_n, 0
            for s in steps:
                acc -= s
                n_fs += 1 << acc
            dev = f.device if _is_torch(f) else torch.device("cuda", self.device)
            fs = torch.empty((n_fs, 8), dtype=torch.int32, device=dev)
        alphas = np.zeros((total, 8), dtype=np.uint32)
        final = np.zeros((1 << (log_n - total), 8), dtype=np.uint32)
        status = capi.lib().zkb_fri_commit_phase(
            self._h, fid, hash_id, log_n, b.ptr, b.mem, (ctypes.c_uint32 * rounds)(*steps), rounds, cb, None, roots,
            trees, fs.data_ptr() if fs is not None else None, capi.u32_ptr(alphas), capi.u32_ptr(final),
            _stream_ptr(f, stream))
        if failure:
            raise failure[0]
        capi.check(status, self._h)
        raw = bytes(roots)
        root_list = [raw[i * db:(i + 1) * db] for i in range(rounds)]
        tree_list = None
        if keep_trees:
            tree_list = [MerkleTree(self, ctypes.c_void_p(trees[i]), root_list[i], db) for i in range(rounds)]
        return {"roots": root_list, "alphas": _ints(alphas), "final_polynomial": _ints(final), "trees": tree_list, "fs": fs}

    # ------------------------------------------------------------------ grinding (proof_of_work.hpp:47-68)
    def pow_grind(self, hash_id, state, mask=0xFFFF, start=0, stream=None):
        """proof_of_work<TranscriptHash, uint32>::generate over a sequential Fiat-Shamir transcript whose current
        digest is `state`: the smallest nonce >= start with (low32(H(H(state || be32(nonce)))) & mask) == 0."""
        db = capi.lib().zkb_merkle_digest_bytes(hash_id)
        state = bytes(state)
        if len(state) != db:
            raise capi.ZkbInvalidArgument(capi.ERR_INVALID_ARGUMENT, "transcript state must be %d bytes" % db)
        buf = (ctypes.c_uint8 * db).from_buffer_copy(state)
        nonce = ctypes.c_uint32(0)
        capi.check(capi.lib().zkb_pow_grind(self._h, hash_id, buf, int(start) & 0xFFFFFFFF, int(mask) & 0xFFFFFFFF,
                                            ctypes.byref(nonce), stream), self._h)
        return int(nonce.value)

    # ------------------------------------------------------------------ LPC opening side (eval_polys, combined Q)
    def poly_evaluate(self, field, polys, n, points, dfs=False, stream=None):
        """polys_evaluator::eval_polys (batched_commitment.hpp:176-190): value of every polynomial of the batch
        ([batch, n, 8]; coefficient form, or evaluations on the 2^k subgroup when dfs) at every point.
        Returns a list (per polynomial) of lists (per point) of integers."""
        fid = _field_id(field)
        b = _Buf(polys)
        batch = b.nbytes // (n * 32)
        pts = np.ascontiguousarray(_int_rows(points))
        out = np.zeros((batch, len(points), 8), dtype=np.uint32)
        capi.check(capi.lib().zkb_poly_evaluate(self._h, fid, capi.POLY_DFS if dfs else capi.POLY_COEFFICIENTS, n, batch,
                                                b.ptr, b.mem, len(points), capi.u32_ptr(pts), capi.u32_ptr(out),
                                                _stream_ptr(polys, stream)), self._h)
        flat = _ints(out.reshape(-1, 8))
        return [flat[i * len(points):(i + 1) * len(points)] for i in range(batch)]

    def poly_evaluate_pm(self, field, polys, n, points, stream=None):
        """Values of coefficient-form device polynomials at z and -z for every point (FRI query phase,
        basic_fri.hpp:819-834).  Returns a list (per polynomial) of lists (per point) of (f(z), f(-z))."""
        fid = _field_id(field)
        b = _Buf(polys)
        if b.mem != capi.MEM_DEVICE:
            raise ValueError("poly_evaluate_pm works on device tensors")
        batch = b.nbytes // (n * 32)
        pts = np.ascontiguousarray(_int_rows(points))
        out = np.zeros((batch, len(points), 2, 8), dtype=np.uint32)
        capi.check(capi.lib().zkb_poly_evaluate_pm(self._h, fid, n, batch, b.ptr, len(points), capi.u32_ptr(pts),
                                                   capi.u32_ptr(out), _stream_ptr(polys, stream)), self._h)
        flat = _ints(out.reshape(-1, 8))
        k = len(points)
        return [[(flat[2 * (i * k + j)], flat[2 * (i * k + j) + 1]) for j in range(k)] for i in range(batch)]

    def poly_lincomb(self, field, polys, n, scalars, constant=None, out=None, accumulate=False, stream=None):
        """out[i] (+)= sum_j scalars[j] polys[j][i] - [i == 0] constant on device tensors (lpc.hpp:139-153)."""
        fid = _field_id(field)
        b = _Buf(polys)
        batch = b.nbytes // (n * 32)
        if len(scalars) != batch:
            raise ValueError("one scalar per polynomial")
        if out is None:
            if accumulate:
                raise ValueError("accumulate needs an existing output")
            out = _empty_like(polys, (n, 8))
        o = _Buf(out, writable=True)
        if b.mem != capi.MEM_DEVICE or o.mem != capi.MEM_DEVICE:
            raise ValueError("poly_lincomb works on device tensors")
        sc = np.ascontiguousarray(_int_rows(scalars))
        capi.check(capi.lib().zkb_poly_lincomb(self._h, fid, n, batch, b.ptr, capi.u32_ptr(sc),
                                               _limbs(constant, 8) if constant is not None else None, o.ptr,
                                               1 if accumulate else 0, _stream_ptr(polys, stream)), self._h)
        return out

    def poly_div_linear(self, field, poly, n, point, out=None, stream=None):
        """(poly - poly(point)) / (X - point) on device tensors; returns (quotient, remainder)."""
        fid = _field_id(field)
        b = _Buf(poly)
        if out is None:
            out = _empty_like(poly, (n, 8))
        o = _Buf(out, writable=True)
        if b.mem != capi.MEM_DEVICE or o.mem != capi.MEM_DEVICE:
            raise ValueError("poly_div_linear works on device tensors")
        rem = np.zeros((1, 8), dtype=np.uint32)
        capi.check(capi.lib().zkb_poly_div_linear(self._h, fid, n, b.ptr, _limbs(point, 8), o.ptr, capi.u32_ptr(rem),
                                                  _stream_ptr(poly, stream)), self._h)
        return out, _ints(rem)[0]

    # ------------------------------------------------------------------ Placeholder permutation argument (8(f)-3)
    def permutation_grand_product(self, field, columns, s_id, s_sigma, beta, gamma, out=None, stream=None):
        """V_P of permutation_argument.hpp:104-133 for device tensors [ncols, n, 8] (columns in global_indices order):
        V_P[0] = 1, V_P[j] = V_P[j-1] prod_i (c_i + beta id_i + gamma)[j-1] / prod_i (c_i + beta sigma_i + gamma)[j-1]."""
        fid = _field_id(field)
        c, a, b = _Buf(columns), _Buf(s_id), _Buf(s_sigma)
        if not (c.mem == a.mem == b.mem == capi.MEM_DEVICE) or not (c.nbytes == a.nbytes == b.nbytes):
            raise ValueError("columns, s_id and s_sigma must be device tensors of the same shape [ncols, n, 8]")
        ncols, n = int(columns.shape[0]), int(columns.shape[1])
        if out is None:
            out = _empty_like(columns, (n, 8))
        o = _Buf(out, writable=True)
        capi.check(capi.lib().zkb_permutation_grand_product(self._h, fid, n, ncols, c.ptr, a.ptr, b.ptr, _limbs(beta, 8),
                                                            _limbs(gamma, 8), o.ptr, _stream_ptr(columns, stream)), self._h)
        return out

    def lookup_grand_product(self, field, inputs, values, sorted_, beta, gamma, usable_rows, out=None, stream=None):
        """compute_V_L (lookup_argument.hpp:375-409) for device tensors reduced_input / reduced_value / sorted of shape
        [count, n, 8]."""
        fid = _field_id(field)
        bufs = [_Buf(t) for t in (inputs, values, sorted_)]
        if any(b.mem != capi.MEM_DEVICE for b in bufs):
            raise ValueError("lookup_grand_product works on device tensors")
        n = int(sorted_.shape[1])
        if out is None:
            out = _empty_like(sorted_, (n, 8))
        o = _Buf(out, writable=True)
        capi.check(capi.lib().zkb_lookup_grand_product(self._h, fid, n, int(usable_rows), int(inputs.shape[0]), bufs[0].ptr,
                                                       int(values.shape[0]), bufs[1].ptr, int(sorted_.shape[0]), bufs[2].ptr,
                                                       _limbs(beta, 8), _limbs(gamma, 8), o.ptr, _stream_ptr(sorted_, stream)),
                   self._h)
        return out

    def prefix_product(self, field, x, exclusive=True, out=None, stream=None):
        """out[i] = prod_{j<i} x[j] (exclusive, out[0] = 1) or prod_{j<=i} x[j] for a device tensor [n, 8]."""
        fid = _field_id(field)
        b = _Buf(x)
        if out is None:
            out = _empty_like(x, tuple(x.shape))
        o = _Buf(out, writable=True)
        if b.mem != capi.MEM_DEVICE or o.mem != capi.MEM_DEVICE:
            raise ValueError("prefix_product works on device tensors")
        capi.check(capi.lib().zkb_prefix_product(self._h, fid, b.nbytes // 32, b.ptr, o.ptr, 1 if exclusive else 0,
                                                 _stream_ptr(x, stream)), self._h)
        return out

    def batch_inverse(self, field, x, out=None, stream=None):
        """out[i] = x[i]^-1 for a device tensor [n, 8]; raises ZkbInvalidArgument if an element is zero."""
        fid = _field_id(field)
        b = _Buf(x)
        if out is None:
            out = _empty_like(x, tuple(x.shape))
        o = _Buf(out, writable=True)
        if b.mem != capi.MEM_DEVICE or o.mem != capi.MEM_DEVICE:
            raise ValueError("batch_inverse works on device tensors")
        capi.check(capi.lib().zkb_batch_inverse(self._h, fid, b.nbytes // 32, b.ptr, o.ptr, _stream_ptr(x, stream)), self._h)
        return out

    # ------------------------------------------------------------------ Placeholder argument builders
    def expr_eval(self, field, cols, program, constants, out, out_stride=1, out_offset=0, accumulate=False, stream=None):
        """zkb_expr_eval: postfix `program` (list of (op, a, b)) over cols [ncols, n, 8] (device) -> out (device)."""
        fid = _field_id(field)
        b = _Buf(cols)
        ncols, n = int(cols.shape[0]), int(cols.shape[1])
        prog = np.zeros((len(program), 3), dtype=np.int32)
        for k, ins in enumerate(program):
            prog[k] = (ins[0], ins[1], ins[2] if len(ins) > 2 else 0)
        cs = _int_rows(list(constants)) if len(constants) else np.zeros((1, 8), dtype=np.uint32)
        o = _Buf(out, writable=True)
        if b.mem != capi.MEM_DEVICE or o.mem != capi.MEM_DEVICE:
            raise ValueError("expr_eval works on device tensors")
        capi.check(capi.lib().zkb_expr_eval(self._h, fid, n, ncols, b.ptr, prog.ctypes.data, len(program), capi.u32_ptr(cs), len(constants),
                                            o.ptr, out_stride, out_offset, int(bool(accumulate)), _stream_ptr(cols, stream)), self._h)
        return out

    def quotient_split(self, field, f_coefficients, log_n, log_ext, nchunks, out=None, stream=None):
        import torch
        fid = _field_id(field)
        b = _Buf(f_coefficients)
        if out is None:
            out = torch.empty((nchunks, 1 << log_n, 8), dtype=torch.int32, device=f_coefficients.device)
        o = _Buf(out, writable=True)
        capi.check(capi.lib().zkb_quotient_split(self._h, fid, log_n, log_ext, b.ptr, nchunks, o.ptr, _stream_ptr(f_coefficients, stream)), self._h)
        return out

    def lookup_sort(self, field, inputs, values, usable_rows, out=None, stream=None):
        """sort_polynomials (lookup_argument.hpp:565-633): inputs [ni, n, 8], values [nv, n, 8] -> sorted [ni + nv, n, 8]"""
        import torch
        fid = _field_id(field)
        ni, nv, n = (int(inputs.shape[0]) if inputs is not None else 0), int(values.shape[0]), int(values.shape[1])
        if out is None:
            out = torch.empty((ni + nv, n, 8), dtype=torch.int32, device=values.device)
        capi.check(capi.lib().zkb_lookup_sort(self._h, fid, n, usable_rows, ni, _Buf(inputs).ptr if ni else None, nv, _Buf(values).ptr,
                                              _Buf(out, writable=True).ptr, _stream_ptr(values, stream)), self._h)
        return out

    # ------------------------------------------------------------------ R1CS rows (r1cs_to_qap.hpp:245-248,289-291)
    def sparse_matrix(self, field, rows, cols, row_ptr, col_idx, values, stream=None):
        return SparseMatrix(self, field, rows, cols, row_ptr, col_idx, values, stream)

    # ------------------------------------------------------------------ MSM
    def msm_bases(self, curve, points, stream=None):
        return MsmBases(self, curve, points, stream)

    def multiexp(self, bases, scalars, offset=0, n=None, stream=None):
        """algebra::multiexp / multiexp_with_mixed_addition over bases[offset:offset+n].
        Returns the affine result as (x, y) Python ints, or None for the point at infinity."""
        sb = _Buf(scalars)
        n = sb.nbytes // 32 if n is None else n
        cl = bases.coord_limbs
        res = (ctypes.c_uint32 * (2 * cl))()
        capi.check(capi.lib().zkb_msm(self._h, bases._h, offset, n, sb.ptr, sb.mem, res, _stream_ptr(scalars, stream)),
                   self._h)
        return _affine_from_limbs(res, cl, bases.deg)

    def multiexp_partial(self, bases, scalars, offset=0, n=None, stream=None):
        """Partial sum in XYZZ/Montgomery limbs (numpy [4*coord_limbs]) for multi-GPU point sharding."""
        sb = _Buf(scalars)
        n = sb.nbytes // 32 if n is None else n
        cl = bases.coord_limbs
        res = np.zeros(4 * cl, dtype=np.uint32)
        capi.check(capi.lib().zkb_msm_partial(self._h, bases._h, offset, n, sb.ptr, sb.mem, capi.u32_ptr(res),
                                              _stream_ptr(scalars, stream)), self._h)
        return res

    def batch_exp(self, curve, base, scalars, stream=None):
        """algebra::batch_exp (generator.hpp:187-225): scalars[i] * base for one affine base (Python ints, (c0, c1) pairs
        on the G2 groups) and a [n, 8] scalar array / device tensor.  Returns affine points [n, 2, coord_limbs] in the
        memory space of `scalars` (zero scalar -> all-zero point)."""
        c = _curve(curve)
        cl = coord_limbs(c)
        parts = []
        for k in range(2):
            comp = base[k] if c.deg == 2 else (base[k],)
            h = cl // len(comp)
            for v in comp:
                parts += [(int(v) >> (32 * l)) & 0xFFFFFFFF for l in range(h)]
        b = (ctypes.c_uint32 * (2 * cl))(*parts)
        sb = _Buf(scalars)
        n = sb.nbytes // 32
        out = _empty_like(scalars, (n, 2, cl))
        o = _Buf(out, writable=True)
        capi.check(capi.lib().zkb_batch_exp(self._h, c.cid, n, b, sb.ptr, o.ptr, sb.mem, _stream_ptr(scalars, stream)), self._h)
        return out

    def points_decompress(self, curve, octets, n, offset=0, stride=None, status=False, stream=None):
        """zkb_points_decompress: n compressed points of the reference's wire format (48-byte G1 / 96-byte G2 encodings,
        `stride` bytes apart, starting at byte `offset` of `octets`: bytes-like, a uint8 numpy array or a uint8 device
        tensor) -> affine points [n, 2, coord_limbs] (device tensor when `octets` is one, else a numpy array).  Raises
        ZkbInvalidArgument when an encoding is not a point (invalid_msg_data upstream); with status=True returns
        (points, status bytes) instead and leaves the judgement to the caller."""
        import torch
        c = _curve(curve)
        cl = coord_limbs(c)
        width = 4 * cl
        stride = width if stride is None else int(stride)
        if _is_torch(octets):
            if octets.dtype != torch.uint8 or not octets.is_cuda or not octets.is_contiguous():
                raise ValueError("octets must be a contiguous uint8 CUDA tensor")
            total, ptr, mem = octets.numel(), octets.data_ptr() + offset, capi.MEM_DEVICE
            out = torch.empty((n, 2, cl), dtype=torch.int32, device=octets.device)
            st = torch.empty((n,), dtype=torch.uint8, device=octets.device) if status else None
            optr, sptr = out.data_ptr(), (st.data_ptr() if status else None)
        else:
            arr = np.frombuffer(octets, dtype=np.uint8) if not isinstance(octets, np.ndarray) else np.ascontiguousarray(octets, dtype=np.uint8)
            total, ptr, mem = arr.size, arr.ctypes.data + offset, capi.MEM_HOST
            out = np.zeros((n, 2, cl), dtype=np.uint32)
            st = np.zeros((n,), dtype=np.uint8) if status else None
            optr, sptr = out.ctypes.data, (st.ctypes.data if status else None)
        if n and offset + (n - 1) * stride + width > total:
            raise ValueError("octets too short for %d points" % n)
        rc = capi.lib().zkb_points_decompress(self._h, c.cid, n, ptr, stride, optr, sptr, mem, _stream_ptr(octets if _is_torch(octets) else None, stream))
        if status and rc == capi.ERR_INVALID_ARGUMENT:
            capi.lib().zkb_ctx_clear_error(self._h)
            return out, st
        capi.check(rc, self._h)
        return (out, st) if status else out

    def grid_points(self, curve, n, table_a, table_b, stream=None):
        """Synthetic bases on the device: out[i] = table_a[i % m] + table_b[i // m] (torch int32 [n,2,cl])."""
        import torch
        c = _curve(curve)
        cl = coord_limbs(c)
        ta = np.ascontiguousarray(table_a, dtype=np.uint32)
        tb = np.ascontiguousarray(table_b, dtype=np.uint32)
        m = ta.size // (2 * cl)
        if tb.size // (2 * cl) < (n + m - 1) // m:
            raise ValueError("table_b too small")
        out = torch.empty((n, 2, cl), dtype=torch.int32, device="cuda:%d" % self.device)
        capi.check(capi.lib().zkb_g1_grid_points(self._h, c.cid, n, m, ta.ctypes.data, tb.ctypes.data, out.data_ptr(),
                                                 _stream_ptr(out, stream)), self._h)
        return out

    def bench_imad_wide(self, blocks, threads=256, iters=4096):
        """bare IMAD.WIDE issue rate (wide multiply-adds/s): the ceiling behind the field-product peaks"""
        r = ctypes.c_double()
        capi.check(capi.lib().zkb_bench_imad_wide(self._h, blocks, threads, iters, ctypes.byref(r)), self._h)
        return r.value

    def bench_field_mul(self, field, blocks, threads=256, iters=4096):
        r = ctypes.c_double()
        capi.check(capi.lib().zkb_bench_field_mul(self._h, _field_id(field), blocks, threads, iters, ctypes.byref(r)), self._h)
        return r.value


class MultiContext:
    """zkb_multi: one process driving several GPUs (include/zkb200.h, multi-GPU section).  Host buffers in, results out."""

    def __init__(self, devices):
        devs = (ctypes.c_int * len(devices))(*[int(d) for d in devices])
        self._h = ctypes.c_void_p()
        capi.check(capi.lib().zkb_multi_create(devs, len(devices), ctypes.byref(self._h)))
        self.devices = list(devices)

    def _check(self, status):
        if status != capi.OK:
            detail = capi.lib().zkb_multi_last_error(self._h).decode()
            try:
                capi.check(status)
            except capi.ZkbError as e:
                raise type(e)(status, str(e) + (": " + detail if detail else "")) from None

    def close(self):
        if self._h:
            capi.lib().zkb_multi_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def msm_bases(self, curve, points_host, precompute=False, max_bytes_per_device=8 << 30):
        c = _curve(curve)
        pts = np.ascontiguousarray(points_host, dtype=np.uint32)
        cl = coord_limbs(c)
        n = pts.size // (2 * cl)
        h = ctypes.c_void_p()
        self._check(capi.lib().zkb_msm_bases_multi_create(self._h, c.cid, n, pts.ctypes.data, ctypes.byref(h)))
        if precompute:
            self._check(capi.lib().zkb_msm_bases_multi_precompute(self._h, h, 0, max_bytes_per_device))
        return _MultiBases(h, c, n)

    def multiexp(self, bases, scalars_host):
        sc = np.ascontiguousarray(scalars_host, dtype=np.uint32)
        cl = coord_limbs(bases.curve)
        res = (ctypes.c_uint32 * (2 * cl))()
        self._check(capi.lib().zkb_msm_multi(self._h, bases._h, sc.size // 8, sc.ctypes.data, res))
        return _affine_from_limbs(res, cl, bases.curve.deg)

    def lpc_commit(self, field, hash_id, polys_host, log_n_in, log_n_out, fri_step):
        a = np.ascontiguousarray(polys_host, dtype=np.uint32)
        batch = a.size // (8 << log_n_in)
        db = capi.lib().zkb_merkle_digest_bytes(hash_id)
        root = (ctypes.c_uint8 * max(db, 1))()
        self._check(capi.lib().zkb_lpc_commit_multi(self._h, _field_id(field), hash_id, log_n_in, log_n_out, fri_step, batch,
                                                    a.ctypes.data, root))
        return bytes(root)


class _MultiBases:
    def __init__(self, h, curve, n):
        self._h, self.curve, self.n = h, curve, n

    def free(self):
        if self._h:
            capi.lib().zkb_msm_bases_multi_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def _affine_from_limbs(res, cl, deg=1):
    """(x, y) as Python ints ((c0, c1) tuples per coordinate on the G2 groups); None = infinity."""
    if deg == 2:
        h = cl // 2
        v = [sum(int(res[k * h + i]) << (32 * i) for i in range(h)) for k in range(4)]
        return None if not any(v) else ((v[0], v[1]), (v[2], v[3]))
    x = sum(int(res[i]) << (32 * i) for i in range(cl))
    y = sum(int(res[cl + i]) << (32 * i) for i in range(cl))
    return None if x == 0 and y == 0 else (x, y)


def msm_combine(curve, partials):
    """Adds per-GPU partial sums (rows of XYZZ limbs) on the host -> affine (x, y) or None."""
    c = _curve(curve)
    cl = coord_limbs(c)
    p = np.ascontiguousarray(np.asarray(partials, dtype=np.uint32).reshape(-1, 4 * cl))
    res = (ctypes.c_uint32 * (2 * cl))()
    capi.check(capi.lib().zkb_msm_combine(c.cid, p.shape[0], capi.u32_ptr(p), res))
    return _affine_from_limbs(res, cl, c.deg)


def field_generator(field):
    n = capi.lib().zkb_field_limbs(_field_id(field))
    out = (ctypes.c_uint32 * n)()
    capi.check(capi.lib().zkb_field_generator(_field_id(field), out))
    return sum(int(out[i]) << (32 * i) for i in range(n))


def unity_root(field, log_n):
    n = capi.lib().zkb_field_limbs(_field_id(field))
    out = (ctypes.c_uint32 * n)()
    capi.check(capi.lib().zkb_field_unity_root(_field_id(field), log_n, out))
    return sum(int(out[i]) << (32 * i) for i in range(n))
