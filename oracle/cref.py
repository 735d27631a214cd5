"""ctypes wrapper of oracle/c/liboracle.so - the C restatement of the reference's CPU algorithms
(TEST INFRASTRUCTURE ONLY; also the timed `cpu_baseline` / `--impl reference` leg of bench.py).
Arrays: numpy uint32 [..., limbs] canonical little-endian (same layout as the product ABI)."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "c", "liboracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        srcs = [os.path.join(HERE, "c", f) for f in ("oracle.c", "field_impl.h", "curve_impl.h")]
        if not os.path.exists(SO) or any(os.path.getmtime(s) > os.path.getmtime(SO) for s in srcs):
            subprocess.check_call(["make", "-s", "-C", os.path.join(HERE, "c")])
        L = ctypes.CDLL(SO)
        vp, i, u32, sz, d = ctypes.c_void_p, ctypes.c_int, ctypes.c_uint32, ctypes.c_size_t, ctypes.c_double
        L.orc_threads_available.restype = i
        L.orc_ntt.restype = d
        L.orc_ntt.argtypes = [i, i, u32, vp, i, vp, i]
        L.orc_lde.restype = d
        L.orc_lde.argtypes = [i, i, i, u32, vp, vp, i]
        L.orc_fri_fold.restype = d
        L.orc_fri_fold.argtypes = [i, i, vp, vp, vp]
        L.orc_hash.restype = None
        L.orc_hash.argtypes = [i, vp, sz, vp]
        L.orc_merkle_commit.restype = d
        L.orc_merkle_commit.argtypes = [i, i, i, u32, vp, vp, vp, i]
        L.orc_lpc_commit.restype = d
        L.orc_lpc_commit.argtypes = [i, i, i, i, i, u32, vp, vp, i, ctypes.POINTER(d)]
        L.orc_msm.restype = d
        L.orc_msm.argtypes = [i, sz, vp, vp, vp, i]
        _lib = L
    return _lib


def threads_available():
    return int(lib().orc_threads_available())


def host_cores():
    """Cores this process may run on (sched affinity).  torchrun exports OMP_NUM_THREADS=1, which makes
    omp_get_max_threads() - threads_available() - answer 1: callers pass this count as `threads` explicitly."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _limbs(v, n=8):
    return np.array([(int(v) >> (32 * k)) & 0xFFFFFFFF for k in range(n)], dtype=np.uint32)


def ntt(fid, data, log_n, inverse=False, shift=None, threads=1):
    """In place on data [batch, 2^log_n, 8]; returns seconds spent in the transforms."""
    assert data.dtype == np.uint32 and data.flags["C_CONTIGUOUS"]
    batch = data.size // (8 << log_n)
    sh = _limbs(shift) if shift is not None else None
    return lib().orc_ntt(fid, log_n, batch, _p(data), int(inverse), _p(sh) if sh is not None else None, threads)


def lde(fid, data, log_in, log_out, threads=1):
    batch = data.size // (8 << log_in)
    out = np.empty((batch, 1 << log_out, 8), dtype=np.uint32)
    t = lib().orc_lde(fid, log_in, log_out, batch, _p(np.ascontiguousarray(data)), _p(out), threads)
    return out, t


def fri_fold(fid, f, log_n, alpha):
    out = np.empty((1 << (log_n - 1), 8), dtype=np.uint32)
    a = _limbs(alpha)
    t = lib().orc_fri_fold(fid, log_n, _p(np.ascontiguousarray(f)), _p(a), _p(out))
    return out, t


def hash_bytes(hid, data: bytes):
    out = (ctypes.c_uint8 * 64)()
    buf = (ctypes.c_uint8 * max(len(data), 1)).from_buffer_copy(data or b"\0")
    lib().orc_hash(hid, buf, len(data), out)
    return bytes(out)[:64 if hid == 2 else 32]


def merkle_commit(hid, evals, log_d, fri_step, threads=1, want_nodes=False):
    batch = evals.size // (8 << log_d)
    dl = 64 if hid == 2 else 32
    root = (ctypes.c_uint8 * dl)()
    leaves = (1 << log_d) >> fri_step
    nodes = np.empty((2 * leaves - 1) * dl, dtype=np.uint8) if want_nodes else None
    t = lib().orc_merkle_commit(hid, log_d, fri_step, batch, _p(np.ascontiguousarray(evals)), root,
                                _p(nodes) if want_nodes else None, threads)
    return (bytes(root), t, nodes) if want_nodes else (bytes(root), t)


def lpc_commit(fid, hid, polys, log_in, log_out, fri_step, threads=1):
    batch = polys.size // (8 << log_in)
    dl = 64 if hid == 2 else 32
    root = (ctypes.c_uint8 * dl)()
    lde_s = ctypes.c_double()
    t = lib().orc_lpc_commit(fid, hid, log_in, log_out, fri_step, batch, _p(np.ascontiguousarray(polys)), root, threads,
                             ctypes.byref(lde_s))
    return bytes(root), t, lde_s.value


def msm(cid, points, scalars, threads=1):
    """points [n, 2, coord_limbs], scalars [n, 8] -> ((x, y) ints or None, seconds)."""
    n = scalars.size // 8
    cl = points.size // (2 * n) if n else (12 if cid == 0 else 8)
    out = np.zeros(2 * cl, dtype=np.uint32)
    t = lib().orc_msm(cid, n, _p(np.ascontiguousarray(points)), _p(np.ascontiguousarray(scalars)), _p(out), threads)
    x = sum(int(out[k]) << (32 * k) for k in range(cl))
    y = sum(int(out[cl + k]) << (32 * k) for k in range(cl))
    return (None if x == 0 and y == 0 else (x, y)), t
