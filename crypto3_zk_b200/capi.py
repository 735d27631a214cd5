"""ctypes binding of libzkb200.so (include/zkb200.h).  This is the only way Python reaches the
kernels: there is no eager/PyTorch/CPU fallback - a missing library or device raises."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
# ZKB200_LIB selects another build of the same library (A/B experiments under profiles/); never a fallback
LIB_PATH = os.environ.get("ZKB200_LIB") or os.path.join(HERE, "libzkb200.so")

OK, ERR_INVALID_ARGUMENT, ERR_DOMAIN_TOO_LARGE, ERR_CUDA, ERR_OOM, ERR_NO_DEVICE, ERR_UNSUPPORTED = range(7)
MEM_HOST, MEM_DEVICE = 0, 1
HASH_KECCAK_256, HASH_SHA2_256, HASH_KECCAK_512 = 0, 1, 2
VEC_MUL, VEC_SUB, VEC_ADD, VEC_MUL_SUB_SCALE = 0, 1, 2, 3

# every symbol include/zkb200.h declares (tests check that the library exports all of them)
SYMBOLS = [
    "zkb_version", "zkb_status_string", "zkb_device_count", "zkb_ctx_create", "zkb_ctx_destroy",
    "zkb_ctx_last_error", "zkb_ctx_set_scratch_limit", "zkb_ctx_release_caches", "zkb_ctx_kernel_launches",
    "zkb_field_limbs", "zkb_field_two_adicity", "zkb_field_generator", "zkb_field_unity_root", "zkb_curve_generator",
    "zkb_ntt", "zkb_lde", "zkb_vec", "zkb_fri_fold", "zkb_lpc_commit", "zkb_merkle_commit",
    "zkb_merkle_digest_bytes", "zkb_merkle_root_of_digests", "zkb_merkle_leaves", "zkb_merkle_path", "zkb_merkle_free",
    "zkb_msm_bases_create", "zkb_msm_bases_free", "zkb_msm_bases_size", "zkb_msm_bases_precompute", "zkb_msm", "zkb_msm_partial",
    "zkb_msm_combine", "zkb_msm_g1", "zkb_g1_grid_points", "zkb_bench_field_mul", "zkb_bench_imad_wide", "zkb_ctx_clear_error", "zkb_msm_window_plan", "zkb_expr_eval", "zkb_quotient_split", "zkb_lookup_sort", "zkb_points_decompress",
    "zkb_multi_create", "zkb_multi_destroy", "zkb_multi_size", "zkb_multi_ctx", "zkb_multi_last_error", "zkb_msm_bases_multi_create",
    "zkb_msm_bases_multi_precompute", "zkb_msm_bases_multi_free", "zkb_msm_multi", "zkb_lpc_commit_multi",
    "zkb_fri_commit_phase", "zkb_poly_evaluate", "zkb_poly_lincomb", "zkb_poly_div_linear",
    "zkb_sparse_matrix_create", "zkb_sparse_matrix_free", "zkb_sparse_matvec", "zkb_pow_grind", "zkb_merkle_paths", "zkb_poly_evaluate_pm",
    "zkb_lde_with_coefficients", "zkb_batch_exp", "zkb_permutation_grand_product", "zkb_lookup_grand_product", "zkb_prefix_product", "zkb_batch_inverse", "zkb_buf_alloc", "zkb_buf_free", "zkb_buf_copy", "zkb_buf_zero", "zkb_gather",
]
POLY_COEFFICIENTS, POLY_DFS = 0, 1
EXPR_PUSH_COL, EXPR_PUSH_CONST, EXPR_ADD, EXPR_SUB, EXPR_MUL, EXPR_NEG = range(6)
# int (*zkb_fri_challenge_fn)(void *user, uint32_t round, const uint8_t *root, uint32_t root_bytes, uint32_t count, uint32_t *alphas_out)
FRI_CHALLENGE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint8),
                                    ctypes.c_uint32, ctypes.c_uint32, ctypes.POINTER(ctypes.c_uint32))


class ZkbError(RuntimeError):
    def __init__(self, status, msg):
        super().__init__("zkb200: %s (status %d)" % (msg, status))
        self.status = status


class ZkbInvalidArgument(ZkbError, ValueError):
    """Mirrors std::invalid_argument of the upstream domain/polynomial classes."""


_lib = None


def lib():
    """Loads libzkb200.so (building it is __graft_entry__.build()'s / crypto3_zk_b200.build's job)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libzkb200.so is missing: run `python -m crypto3_zk_b200.build` (needs nvcc); "
                          "there is no fallback implementation")
    L = ctypes.CDLL(LIB_PATH)
    vp, u32p, u8p = ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_uint8)
    i, u32, u64 = ctypes.c_int, ctypes.c_uint32, ctypes.c_uint64
    L.zkb_version.restype = ctypes.c_char_p
    L.zkb_status_string.restype = ctypes.c_char_p
    L.zkb_status_string.argtypes = [i]
    L.zkb_device_count.restype = i
    L.zkb_ctx_create.argtypes = [i, ctypes.POINTER(vp)]
    L.zkb_ctx_destroy.argtypes = [vp]
    L.zkb_ctx_destroy.restype = None
    L.zkb_ctx_last_error.argtypes = [vp]
    L.zkb_ctx_last_error.restype = ctypes.c_char_p
    L.zkb_ctx_set_scratch_limit.argtypes = [vp, u64]
    L.zkb_ctx_release_caches.argtypes = [vp]
    L.zkb_ctx_kernel_launches.argtypes = [vp]
    L.zkb_ctx_kernel_launches.restype = u64
    L.zkb_field_limbs.argtypes = [i]
    L.zkb_field_two_adicity.argtypes = [i]
    L.zkb_field_generator.argtypes = [i, u32p]
    L.zkb_field_unity_root.argtypes = [i, i, u32p]
    L.zkb_curve_generator.argtypes = [i, u32p]
    L.zkb_ntt.argtypes = [vp, i, i, u32, vp, vp, i, u32p, i, vp]
    L.zkb_lde.argtypes = [vp, i, i, i, u32, vp, vp, i, vp]
    L.zkb_vec.argtypes = [vp, i, i, u64, vp, vp, vp, u32p, vp, i, vp]
    L.zkb_fri_fold.argtypes = [vp, i, i, vp, u32p, vp, i, vp]
    L.zkb_lpc_commit.argtypes = [vp, i, i, i, i, i, u32, vp, i, u8p, ctypes.POINTER(vp), vp]
    L.zkb_merkle_commit.argtypes = [vp, i, i, i, i, u32, vp, i, u8p, ctypes.POINTER(vp), vp]
    L.zkb_merkle_digest_bytes.argtypes = [i]
    L.zkb_merkle_leaves.argtypes = [vp]
    L.zkb_merkle_leaves.restype = u64
    L.zkb_merkle_path.argtypes = [vp, vp, u64, u8p]
    L.zkb_merkle_free.argtypes = [vp]
    L.zkb_merkle_free.restype = None
    L.zkb_msm_bases_create.argtypes = [vp, i, u64, vp, i, vp, ctypes.POINTER(vp)]
    L.zkb_msm_bases_free.argtypes = [vp]
    L.zkb_msm_bases_free.restype = None
    L.zkb_merkle_root_of_digests.argtypes = [vp, i, u32, ctypes.c_char_p, u8p, vp]
    L.zkb_msm_bases_precompute.argtypes = [vp, vp, i, u64, vp]
    L.zkb_msm_bases_size.argtypes = [vp]
    L.zkb_msm_bases_size.restype = u64
    L.zkb_msm.argtypes = [vp, vp, u64, u64, vp, i, u32p, vp]
    L.zkb_msm_partial.argtypes = [vp, vp, u64, u64, vp, i, u32p, vp]
    L.zkb_msm_combine.argtypes = [i, u32, u32p, u32p]
    L.zkb_msm_window_plan.argtypes = [vp, u64, ctypes.POINTER(i), ctypes.POINTER(i), ctypes.POINTER(i)]
    L.zkb_msm_g1.argtypes = [vp, i, u64, vp, vp, i, u32p, vp]
    L.zkb_g1_grid_points.argtypes = [vp, i, u64, u32, vp, vp, vp, vp]
    L.zkb_bench_field_mul.argtypes = [vp, i, u32, u32, u32, ctypes.POINTER(ctypes.c_double)]
    L.zkb_bench_imad_wide.argtypes = [vp, u32, u32, u32, ctypes.POINTER(ctypes.c_double)]
    L.zkb_ctx_clear_error.argtypes = [vp]
    L.zkb_ctx_clear_error.restype = None
    L.zkb_expr_eval.argtypes = [vp, i, u64, u32, vp, vp, u32, u32p, u32, vp, u64, u64, i, vp]
    L.zkb_quotient_split.argtypes = [vp, i, i, i, vp, u32, vp, vp]
    L.zkb_lookup_sort.argtypes = [vp, i, u64, u64, u32, vp, u32, vp, vp, vp]
    L.zkb_points_decompress.argtypes = [vp, i, u64, vp, u64, vp, vp, i, vp]
    L.zkb_multi_create.argtypes = [ctypes.POINTER(i), u32, ctypes.POINTER(vp)]
    L.zkb_multi_destroy.argtypes = [vp]
    L.zkb_multi_destroy.restype = None
    L.zkb_multi_size.argtypes = [vp]
    L.zkb_multi_size.restype = u32
    L.zkb_multi_ctx.argtypes = [vp, u32]
    L.zkb_multi_ctx.restype = vp
    L.zkb_multi_last_error.argtypes = [vp]
    L.zkb_multi_last_error.restype = ctypes.c_char_p
    L.zkb_msm_bases_multi_create.argtypes = [vp, i, u64, vp, ctypes.POINTER(vp)]
    L.zkb_msm_bases_multi_precompute.argtypes = [vp, vp, i, u64]
    L.zkb_msm_bases_multi_free.argtypes = [vp]
    L.zkb_msm_bases_multi_free.restype = None
    L.zkb_msm_multi.argtypes = [vp, vp, u64, vp, u32p]
    L.zkb_lpc_commit_multi.argtypes = [vp, i, i, i, i, i, u32, vp, u8p]
    L.zkb_fri_commit_phase.argtypes = [vp, i, i, i, vp, i, u32p, u32, FRI_CHALLENGE_FN, vp, u8p, ctypes.POINTER(vp), vp,
                                       u32p, u32p, vp]
    L.zkb_poly_evaluate.argtypes = [vp, i, i, u64, u32, vp, i, u32, u32p, u32p, vp]
    L.zkb_poly_evaluate_pm.argtypes = [vp, i, u64, u32, vp, u32, u32p, u32p, vp]
    L.zkb_poly_lincomb.argtypes = [vp, i, u64, u32, vp, u32p, u32p, vp, i, vp]
    L.zkb_poly_div_linear.argtypes = [vp, i, u64, vp, u32p, vp, u32p, vp]
    L.zkb_sparse_matrix_create.argtypes = [vp, i, u64, u64, ctypes.POINTER(u64), u32p, u32p, vp, ctypes.POINTER(vp)]
    L.zkb_sparse_matrix_free.argtypes = [vp]
    L.zkb_sparse_matrix_free.restype = None
    L.zkb_sparse_matvec.argtypes = [vp, vp, vp, i, vp, vp]
    L.zkb_merkle_paths.argtypes = [vp, vp, u32, ctypes.POINTER(u64), u8p, vp]
    L.zkb_lde_with_coefficients.argtypes = [vp, i, i, i, u32, vp, vp, vp, vp]
    L.zkb_batch_exp.argtypes = [vp, i, u64, u32p, vp, vp, i, vp]
    L.zkb_permutation_grand_product.argtypes = [vp, i, u64, u32, vp, vp, vp, u32p, u32p, vp, vp]
    L.zkb_lookup_grand_product.argtypes = [vp, i, u64, u64, u32, vp, u32, vp, u32, vp, u32p, u32p, vp, vp]
    L.zkb_prefix_product.argtypes = [vp, i, u64, vp, vp, i, vp]
    L.zkb_batch_inverse.argtypes = [vp, i, u64, vp, vp, vp]
    L.zkb_buf_alloc.argtypes = [vp, u64, ctypes.POINTER(vp)]
    L.zkb_buf_free.argtypes = [vp, vp]
    L.zkb_buf_free.restype = None
    L.zkb_buf_copy.argtypes = [vp, vp, i, vp, i, u64, vp]
    L.zkb_buf_zero.argtypes = [vp, vp, u64, vp]
    L.zkb_gather.argtypes = [vp, vp, u32, ctypes.POINTER(u64), u32p, vp]
    L.zkb_pow_grind.argtypes = [vp, i, u8p, u32, u32, u32p, vp]
    _lib = L
    return L


def check(status, ctx=None):
    if status == OK:
        return
    L = lib()
    msg = L.zkb_status_string(status).decode()
    if ctx:
        detail = L.zkb_ctx_last_error(ctx).decode()
        L.zkb_ctx_clear_error(ctx)      # consumed: a later failure that records no message must not show this one
        if detail:
            msg += ": " + detail
    if status in (ERR_INVALID_ARGUMENT, ERR_DOMAIN_TOO_LARGE):
        raise ZkbInvalidArgument(status, msg)
    raise ZkbError(status, msg)


def u32_ptr(arr):
    """ctypes uint32* view of a contiguous numpy uint32 array."""
    return arr.ctypes.data_as(ctypes.POINTER(ctypes.c_uint32))
