"""ctypes loader of libcholesky_b200.so (the C ABI of include/cholesky.h).  Fails loudly when the
library is missing: there is no Python or CPU fallback for the numeric path."""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcholesky_b200.so")

# every symbol include/cholesky.h, include/chol_mmio.h and include/chol_mnd.h declare
EXPORTS = [
    "chol_create", "chol_num_ranks", "chol_rank_handle", "chol_destroy", "chol_last_error", "register_mappers", "chol_load", "chol_load_arrays",
    "chol_generate", "chol_write_inputs", "chol_analyze", "chol_save_analysis", "chol_load_analysis", "chol_n", "chol_nz", "chol_levels",
    "chol_num_separators", "chol_max_int_size", "chol_num_blocks", "chol_num_clusters0", "chol_get_perm",
    "chol_get_sep_sizes", "chol_get_block_bounds", "chol_num_filled", "chol_get_filled", "chol_filled_checksum",
    "chol_flops", "chol_flops_by_level", "chol_call_counts", "chol_factor_doubles", "chol_level_bytes", "chol_assemble", "chol_factor",
    "chol_fused_dpotrf", "chol_fused_dtrsm", "chol_fused_update", "chol_factor_host", "chol_synchronize",
    "chol_kernel_times", "chol_launch_times", "chol_num_launches", "chol_get_launch", "chol_set_partition", "chol_ipc_export",
    "chol_ipc_import", "chol_partition_stats", "chol_rank", "chol_world", "chol_top_copies_diff", "chol_factor_nnz", "chol_get_factor_coo", "chol_get_factor_dense", "chol_write_factor", "chol_write_factor_binary", "chol_factor_binary_to_mtx",
    "chol_residual", "chol_residual_partial", "chol_residual_finish", "chol_write_debug_log", "chol_factor_debug", "chol_solve", "chol_solve_top_size", "chol_solve_forward", "chol_solve_backward", "chol_solve_stats", "chol_matvec", "chol_read_vector", "chol_write_solution",
    "mm_read_banner", "mm_read_mtx_crd_size", "mm_write_banner", "mm_write_mtx_crd_size", "mm_typecode_to_str",
    "mnd_read_separators", "mnd_read_clusters", "mnd_read_matrix", "mnd_read_vector", "mnd_hash_sax",
]


class Stats(C.Structure):
    _fields_ = [("seconds_best", C.c_double), ("seconds_median", C.c_double), ("seconds_last", C.c_double),
                ("assemble_seconds", C.c_double), ("flops", C.c_double), ("kernel_launches", C.c_int64),
                ("info", C.c_int)]


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(make -C cholesky_b200/csrc). The engine has no fallback path.")
    L = C.CDLL(LIB_PATH)
    L.chol_last_error.restype = C.c_char_p
    L.chol_rank_handle.restype = C.c_void_p
    L.chol_flops.restype = C.c_double
    L.chol_filled_checksum.restype = C.c_uint64
    L.mnd_hash_sax.restype = C.c_uint64
    L.mnd_hash_sax.argtypes = [C.c_uint64]
    for name in ("chol_nz", "chol_num_blocks", "chol_num_clusters0", "chol_get_block_bounds", "chol_num_filled",
                 "chol_get_filled", "chol_factor_doubles", "chol_solve_top_size", "chol_num_launches", "chol_launch_times", "chol_factor_nnz", "chol_get_factor_coo",
                 "mnd_read_clusters"):
        getattr(L, name).restype = C.c_int64
    _lib = L
    return L
