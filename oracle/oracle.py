"""ctypes wrapper of the CPU oracle (oracle/chol_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under cholesky_b200/ may import this.
"""
import ctypes as C
import glob
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libchol_oracle.so")


class Filled(C.Structure):
    """reference `fspace Filled` (blas.rg:55-61): 9 x int64, filled == 0 means FILLED"""
    _fields_ = [(k, C.c_int64) for k in
                ("filled", "sep_x", "sep_y", "interval", "cluster", "lo_x", "lo_y", "hi_x", "hi_y")]


def build(force=False):
    src = os.path.join(HERE, "chol_oracle.c")
    if force or not os.path.exists(LIB) or os.path.getmtime(LIB) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", HERE, "-s"])


def find_openblas():
    """the LP64 OpenBLAS bundled with scipy (`scipy_` symbol prefix)"""
    env = os.environ.get("CHOL_ORACLE_BLAS")
    if env:
        return env
    for p in sys.path:
        hits = sorted(glob.glob(os.path.join(p, "scipy.libs", "libscipy_openblas*.so")))
        if hits:
            return hits[0]
    import scipy
    hits = sorted(glob.glob(os.path.join(os.path.dirname(os.path.dirname(scipy.__file__)),
                                         "scipy.libs", "libscipy_openblas*.so")))
    if hits:
        return hits[0]
    raise RuntimeError("no host OpenBLAS found for the oracle")


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(LIB)
    P = C.c_void_p
    L.orc_create.restype = P
    L.orc_last_error.restype = C.c_char_p
    L.orc_blas_config.restype = C.c_char_p
    L.orc_flops.restype = C.c_double
    L.orc_hash_sax.restype = C.c_uint64
    L.orc_hash_sax.argtypes = [C.c_uint64]
    L.orc_filled_checksum.restype = C.c_uint64
    for name in ("orc_num_blocks", "orc_num_clusters0", "orc_get_block_bounds", "orc_num_filled",
                 "orc_get_filled", "orc_factor_nnz_alloc", "orc_factor_nnz", "orc_get_factor_coo"):
        getattr(L, name).restype = C.c_int64
    for name in ("orc_destroy", "orc_last_error", "orc_load", "orc_analyze", "orc_n", "orc_nz", "orc_levels",
                 "orc_num_separators", "orc_max_int_size", "orc_num_blocks", "orc_num_clusters0", "orc_get_perm",
                 "orc_get_sep_sizes", "orc_get_block_bounds", "orc_num_filled", "orc_get_filled",
                 "orc_filled_checksum", "orc_factor_nnz_alloc", "orc_flops", "orc_flops_by_level",
                 "orc_call_counts", "orc_assemble", "orc_factor", "orc_factor_levels", "orc_fused_dpotrf",
                 "orc_fused_dtrsm", "orc_fused_update", "orc_factor_nnz", "orc_get_factor_coo",
                 "orc_get_factor_dense", "orc_write_factor", "orc_solve", "orc_debug_trace"):
        getattr(L, name).argtypes = None  # first arg is the handle; set per call below
    if L.orc_set_blas(find_openblas().encode()) != 0:
        raise RuntimeError("oracle: cannot load host BLAS")
    _lib = L
    return L


def hash_sax(key):
    return int(lib().orc_hash_sax(C.c_uint64(key)))


class Oracle:
    def __init__(self, mtx, ord_file, clust_file, literal_assembly=-1):
        self.L = lib()
        self.h = C.c_void_p(self.L.orc_create())
        if self.L.orc_load(self.h, mtx.encode(), ord_file.encode(), clust_file.encode()) != 0:
            raise RuntimeError("oracle load: " + self.err())
        if self.L.orc_analyze(self.h, C.c_int(literal_assembly)) != 0:
            raise RuntimeError("oracle analyze: " + self.err())
        self.n = self.L.orc_n(self.h)
        self.nz = self.L.orc_nz(self.h)
        self.levels = self.L.orc_levels(self.h)
        self.num_separators = self.L.orc_num_separators(self.h)

    def err(self):
        return self.L.orc_last_error(self.h).decode()

    def close(self):
        if self.h:
            self.L.orc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- structure
    def perm(self):
        out = np.zeros(self.n, dtype=np.int32)
        self.L.orc_get_perm(self.h, out.ctypes.data_as(C.c_void_p))
        return out

    def sep_sizes(self):
        out = np.zeros(self.num_separators, dtype=np.int32)
        self.L.orc_get_sep_sizes(self.h, out.ctypes.data_as(C.c_void_p))
        return out

    def block_bounds(self):
        k = self.L.orc_get_block_bounds(self.h, None)
        out = np.zeros((k, 6), dtype=np.int64)
        self.L.orc_get_block_bounds(self.h, out.ctypes.data_as(C.c_void_p))
        return out

    def num_blocks(self):
        return int(self.L.orc_num_blocks(self.h))

    def num_clusters0(self):
        return int(self.L.orc_num_clusters0(self.h))

    def max_int_size(self):
        return int(self.L.orc_max_int_size(self.h))

    def num_filled(self, lbl):
        return int(self.L.orc_num_filled(self.h, C.c_int(lbl)))

    def filled(self, lbl):
        """(k, 9) int64 array of `Filled` records of interval label lbl, sorted (sep_x, sep_y, cluster)"""
        k = self.num_filled(lbl)
        out = np.zeros((max(k, 1), 9), dtype=np.int64)
        self.L.orc_get_filled(self.h, C.c_int(lbl), out.ctypes.data_as(C.c_void_p))
        return out[:k]

    def filled_checksum(self, lbl):
        return int(self.L.orc_filled_checksum(self.h, C.c_int(lbl)))

    def alloc_doubles(self):
        return int(self.L.orc_factor_nnz_alloc(self.h))

    def flops(self):
        return float(self.L.orc_flops(self.h))

    def flops_by_level(self):
        a = [np.zeros(self.levels) for _ in range(4)]
        self.L.orc_flops_by_level(self.h, *[x.ctypes.data_as(C.c_void_p) for x in a])
        return dict(potrf=a[0], trsm=a[1], syrk=a[2], gemm=a[3])

    def call_counts(self):
        c = np.zeros(4, dtype=np.int64)
        self.L.orc_call_counts(self.h, c.ctypes.data_as(C.c_void_p))
        return dict(potrf=int(c[0]), trsm=int(c[1]), syrk=int(c[2]), gemm=int(c[3]))

    # -- numeric
    def assemble(self):
        if self.L.orc_assemble(self.h) != 0:
            raise RuntimeError(self.err())

    def factor(self, threads=1):
        s = C.c_double(0)
        if self.L.orc_factor(self.h, C.c_int(threads), C.byref(s)) != 0:
            raise RuntimeError(self.err())
        return s.value

    def factor_levels(self, from_level, to_level, threads=1):
        s = C.c_double(0)
        if self.L.orc_factor_levels(self.h, C.c_int(threads), C.c_int(from_level), C.c_int(to_level),
                                    C.byref(s)) != 0:
            raise RuntimeError(self.err())
        return s.value

    def fused_dpotrf(self, lvl):
        assert self.L.orc_fused_dpotrf(self.h, C.c_int(lvl)) == 0

    def fused_dtrsm(self, lvl):
        assert self.L.orc_fused_dtrsm(self.h, C.c_int(lvl)) == 0

    def fused_update(self, lvl):
        assert self.L.orc_fused_update(self.h, C.c_int(lvl)) == 0

    def factor_nnz(self):
        return int(self.L.orc_factor_nnz(self.h))

    def factor_coo(self):
        k = self.factor_nnz()
        I = np.zeros(k, dtype=np.int32)
        J = np.zeros(k, dtype=np.int32)
        V = np.zeros(k, dtype=np.float64)
        self.L.orc_get_factor_coo(self.h, I.ctypes.data_as(C.c_void_p), J.ctypes.data_as(C.c_void_p),
                                  V.ctypes.data_as(C.c_void_p))
        return I, J, V

    def factor_dense(self):
        out = np.zeros((self.n, self.n), dtype=np.float64)
        self.L.orc_get_factor_dense(self.h, out.ctypes.data_as(C.c_void_p))
        return out

    def write_factor(self, path, full_precision=False):
        if self.L.orc_write_factor(self.h, path.encode(), C.c_int(1 if full_precision else 0)) != 0:
            raise RuntimeError(self.err())

    def debug_trace(self, directory, log_path, full_precision=False):
        """the reference's `-d` run: log + one snapshot per fused task (verify.debug_factor's input)"""
        os.makedirs(directory, exist_ok=True)
        if self.L.orc_debug_trace(self.h, directory.encode(), log_path.encode(),
                                  C.c_int(1 if full_precision else 0)) != 0:
            raise RuntimeError(self.err())

    def solve(self, b):
        b = np.ascontiguousarray(b, dtype=np.float64).reshape(-1)
        x = np.zeros(self.n, dtype=np.float64)
        if self.L.orc_solve(self.h, b.ctypes.data_as(C.c_void_p), x.ctypes.data_as(C.c_void_p)) != 0:
            raise RuntimeError(self.err())
        return x


def read_vector(path, n):
    out = np.zeros(n, dtype=np.float64)
    if lib().orc_read_vector(path.encode(), C.c_int(n), out.ctypes.data_as(C.c_void_p)) != 0:
        raise RuntimeError("oracle read_vector failed: " + path)
    return out


def write_solution(path, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    assert lib().orc_write_solution(path.encode(), C.c_int(x.size), x.ctypes.data_as(C.c_void_p)) == 0


def blas_config():
    return lib().orc_blas_config().decode()
