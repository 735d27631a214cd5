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

    def free(self):
        if self._h:
            capi.lib().zkb_merkle_free(self._h)
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

    def lde(self, field, data, log_n_in, log_n_out, out=None, stream=None):
        """polynomial_dfs::resize(2^log_n_out) for data[batch, 2^log_n_in, 8] -> [batch, 2^log_n_out, 8]."""
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

    def bench_field_mul(self, field, blocks, threads=256, iters=4096):
        r = ctypes.c_double()
        capi.check(capi.lib().zkb_bench_field_mul(self._h, _field_id(field), blocks, threads, iters, ctypes.byref(r)), self._h)
        return r.value


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
