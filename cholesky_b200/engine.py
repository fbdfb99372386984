"""Host-side mirror of the reference's program for the factorization path (mmat.rg main, 1056-1362):
load (-i/-s/-c), symbolic analysis, numeric factorization on the GPU, factor output (-m).  Everything
numeric happens inside libcholesky_b200.so (CUDA, sm_100a)."""
import ctypes as C

import numpy as np

from . import _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class CholeskyError(RuntimeError):
    pass


class Cholesky:
    def __init__(self, device=0, devices=None):
        """one GPU (`device`), or a group of 2, 4 or 8 GPUs driven from this process (`devices`: CUDA ordinals;
        repeating one makes several ranks share it)"""
        self.L = _lib.load()
        self.h = C.c_void_p()
        devs = list(devices) if devices is not None else [device]
        arr = (C.c_int * len(devs))(*devs)
        if self.L.chol_create(arr, len(devs), C.byref(self.h)) != 0:
            raise CholeskyError(f"chol_create failed (devices {devs}: 1, 2, 4 or 8 GPUs)")
        self._owned = True

    @classmethod
    def _borrowed(cls, handle, parent):
        self = cls.__new__(cls)
        self.L, self.h, self._owned, self._parent = parent.L, C.c_void_p(handle), False, parent
        for k in ("n", "nz", "levels", "num_separators"):
            if hasattr(parent, k):
                setattr(self, k, getattr(parent, k))
        return self

    def num_ranks(self):
        return int(self.L.chol_num_ranks(self.h))

    def rank_handle(self, r):
        """rank r of a group handle, for the per-rank inspection calls (partition_stats, launches, ...)"""
        h = self.L.chol_rank_handle(self.h, r)
        if not h:
            raise CholeskyError(f"no rank {r}")
        return Cholesky._borrowed(h, self)

    # ---- lifecycle
    def close(self):
        if getattr(self, "h", None) and getattr(self, "_owned", False):
            self.L.chol_destroy(self.h)
        self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc != 0:
            raise CholeskyError(self.L.chol_last_error(self.h).decode())

    # ---- inputs (mmat.rg:1097-1130)
    def load(self, matrix, separators, clusters):
        self._ck(self.L.chol_load(self.h, matrix.encode(), separators.encode(), clusters.encode()))
        return self

    def generate(self, nx, ny=1, nz=1, stencil=7, levels=0):
        self._ck(self.L.chol_generate(self.h, nx, ny, nz, stencil, levels))
        return self

    def write_inputs(self, matrix=None, separators=None, clusters=None):
        def enc(s):
            return s.encode() if s else None
        self._ck(self.L.chol_write_inputs(self.h, enc(matrix), enc(separators), enc(clusters)))

    # ---- symbolic (mmat.rg:1134-1209)
    def analyze(self, keep_records=False):
        self._ck(self.L.chol_analyze(self.h, 1 if keep_records else 0))
        self.n = self.L.chol_n(self.h)
        self.nz = int(self.L.chol_nz(self.h))
        self.levels = self.L.chol_levels(self.h)
        self.num_separators = self.L.chol_num_separators(self.h)
        return self

    def _after_analyze(self):
        self.n = self.L.chol_n(self.h)
        self.nz = int(self.L.chol_nz(self.h))
        self.levels = self.L.chol_levels(self.h)
        self.num_separators = self.L.chol_num_separators(self.h)
        return self

    def save_analysis(self, path):
        """the symbolic analysis as a file (the other ranks of a node load it instead of analysing again)"""
        self._ck(self.L.chol_save_analysis(self.h, path.encode()))

    def load_analysis(self, path):
        self._ck(self.L.chol_load_analysis(self.h, path.encode()))
        return self._after_analyze()

    def perm(self):
        out = np.zeros(self.n, dtype=np.int32)
        self.L.chol_get_perm(self.h, _p(out))
        return out

    def sep_sizes(self):
        out = np.zeros(self.num_separators, dtype=np.int32)
        self.L.chol_get_sep_sizes(self.h, _p(out))
        return out

    def block_bounds(self):
        k = self.L.chol_get_block_bounds(self.h, None)
        out = np.zeros((k, 6), dtype=np.int64)
        self.L.chol_get_block_bounds(self.h, _p(out))
        return out

    def num_blocks(self):
        return int(self.L.chol_num_blocks(self.h))

    def num_clusters0(self):
        return int(self.L.chol_num_clusters0(self.h))

    def max_int_size(self):
        return int(self.L.chol_max_int_size(self.h))

    def num_filled(self, lbl):
        return int(self.L.chol_num_filled(self.h, lbl))

    def filled(self, lbl):
        k = self.num_filled(lbl)
        out = np.zeros((max(k, 1), 9), dtype=np.int64)
        got = self.L.chol_get_filled(self.h, lbl, _p(out))
        if got < 0:
            raise CholeskyError(self.L.chol_last_error(self.h).decode())
        return out[:k]

    def filled_checksum(self, lbl):
        return int(self.L.chol_filled_checksum(self.h, lbl))

    def flops(self):
        return float(self.L.chol_flops(self.h))

    def flops_by_level(self):
        a = [np.zeros(self.levels) for _ in range(4)]
        self.L.chol_flops_by_level(self.h, *[_p(x) for x in a])
        return dict(potrf=a[0], trsm=a[1], syrk=a[2], gemm=a[3])

    def call_counts(self):
        c = np.zeros(4, dtype=np.int64)
        self.L.chol_call_counts(self.h, _p(c))
        return dict(potrf=int(c[0]), trsm=int(c[1]), syrk=int(c[2]), gemm=int(c[3]))

    def level_bytes(self, lvl):
        """algorithmic HBM bytes of a tree level: panels (in-place factorization), Schur operands, destinations"""
        out = np.zeros(3, dtype=np.float64)
        self._ck(self.L.chol_level_bytes(self.h, lvl, _p(out)))
        return dict(panel=float(out[0]), operands=float(out[1]), destinations=float(out[2]))

    def factor_doubles(self):
        return int(self.L.chol_factor_doubles(self.h))

    # ---- numeric (mmat.rg:1211-1358), GPU only
    def assemble(self):
        self._ck(self.L.chol_assemble(self.h))

    def factor(self, iterations=1, warmup=0):
        st = _lib.Stats()
        self._ck(self.L.chol_factor(self.h, iterations, warmup, C.byref(st)))
        return st

    def fused_dpotrf(self, lvl):
        self._ck(self.L.chol_fused_dpotrf(self.h, lvl))

    def fused_dtrsm(self, lvl):
        self._ck(self.L.chol_fused_dtrsm(self.h, lvl))

    def fused_update(self, lvl):
        self._ck(self.L.chol_fused_update(self.h, lvl))

    def factor_host(self, values=None):
        """end to end with host buffers: H2D of A's values, assemble, factor, D2H of diag(L)"""
        st = _lib.Stats()
        diag = np.zeros(self.n, dtype=np.float64)
        if values is None:
            self._ck(self.L.chol_factor_host(self.h, None, C.c_int64(0), _p(diag), C.byref(st)))
        else:
            v = np.ascontiguousarray(values, dtype=np.float64)
            self._ck(self.L.chol_factor_host(self.h, _p(v), C.c_int64(v.size), _p(diag), C.byref(st)))
        return diag, st

    def kernel_times(self):
        a, b, c, f = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        self._ck(self.L.chol_kernel_times(self.h, C.byref(a), C.byref(b), C.byref(c), C.byref(f)))
        return dict(panel_ms=a.value, exchange_ms=b.value, gemm_ms=c.value, gemm_flops=f.value)

    # ---- multi-GPU: one process and one handle per GPU (see cholesky_b200/distributed.py)
    def set_partition(self, rank, world):
        self._ck(self.L.chol_set_partition(self.h, rank, world))
        return self

    def ipc_export(self):
        buf = (C.c_ubyte * 128)()
        self._ck(self.L.chol_ipc_export(self.h, buf))
        return bytes(buf)

    def ipc_import(self, blobs):
        world = len(blobs)
        raw = b"".join(blobs)
        buf = (C.c_ubyte * len(raw)).from_buffer_copy(raw)
        self._ck(self.L.chol_ipc_import(self.h, buf, world))

    def partition_stats(self):
        out = np.zeros(8, dtype=np.float64)
        self._ck(self.L.chol_partition_stats(self.h, _p(out)))
        return dict(assembled=int(out[0]), gemm_flops=float(out[1]), push_launches=int(out[2]),
                    top_doubles=int(out[3]), diag_tiles=int(out[4]), row_slabs=int(out[5]),
                    reduce_pull_bytes=float(out[6]), push_bytes=float(out[7]))

    def top_copies_diff(self):
        """largest |difference| between the ranks' copies of the factored top panels (0 = bit-identical)"""
        d = C.c_double()
        self._ck(self.L.chol_top_copies_diff(self.h, C.byref(d)))
        return d.value

    def launch_times(self):
        """per-launch device ms of the last kernel_times() pass (launch-list order)"""
        n = int(self.L.chol_num_launches(self.h))
        out = np.zeros(n, dtype=np.float32)
        self.L.chol_launch_times(self.h, _p(out), C.c_int64(n))
        return out

    def launches(self):
        """the compiled launch list as dicts (kind, level, phase, ctas, flops, cfg)"""
        out = []
        kind, level, phase, cfg = C.c_int(), C.c_int(), C.c_int(), C.c_int()
        ctas, flops = C.c_int64(), C.c_double()
        for i in range(int(self.L.chol_num_launches(self.h))):
            self.L.chol_get_launch(self.h, C.c_int64(i), C.byref(kind), C.byref(level), C.byref(phase),
                                   C.byref(ctas), C.byref(flops), C.byref(cfg))
            names = ("panel_kernel", "trsm_tile", "gemm_grouped", "peer_sync", "reduce_rects", "nop", "push_rects")
            out.append(dict(kind=names[kind.value], level=level.value, phase=phase.value, ctas=ctas.value,
                            flops=flops.value, cfg=cfg.value & 15, stream=(cfg.value >> 4) & 15, width=cfg.value >> 8))
        return out

    # ---- results (mmat.rg:1360-1362)
    def factor_nnz(self):
        k = int(self.L.chol_factor_nnz(self.h))
        if k < 0:
            raise CholeskyError(self.L.chol_last_error(self.h).decode())
        return k

    def factor_coo(self):
        k = self.factor_nnz()
        I = np.zeros(k, dtype=np.int32)
        J = np.zeros(k, dtype=np.int32)
        V = np.zeros(k, dtype=np.float64)
        self.L.chol_get_factor_coo(self.h, _p(I), _p(J), _p(V))
        return I, J, V

    def factor_dense(self):
        out = np.zeros((self.n, self.n), dtype=np.float64)
        self._ck(self.L.chol_get_factor_dense(self.h, _p(out)))
        return out

    def write_factor(self, path, full_precision=False):
        self._ck(self.L.chol_write_factor(self.h, path.encode(), 1 if full_precision else 0))

    def write_factor_binary(self, path):
        """binary block dump of the factor (csrc/factor_file.cc); `factor_binary_to_mtx` converts it"""
        self._ck(self.L.chol_write_factor_binary(self.h, path.encode()))

    # ---- debug trace, the `-d` path (mmat.rg:1086-1090; verify.py:216-275 replays it)
    def write_debug_log(self, path=None):
        """the log lines of a debug run (host only; needs analyze(keep_records=True)); None: stdout"""
        self._ck(self.L.chol_write_debug_log(self.h, path.encode() if path else None))

    def factor_debug(self, directory, full_precision=False, with_txt=False):
        """level loop one fused task group at a time on the GPU, one snapshot file per reference task"""
        self._ck(self.L.chol_factor_debug(self.h, directory.encode(), 1 if full_precision else 0, 1 if with_txt else 0))

    def residual(self, k=4, seed=1):
        """randomized ||(A - L L^T) W|| / ||A W||, k <= 4 probe columns, evaluated on the GPU (single-GPU or group handle)"""
        r = C.c_double()
        self._ck(self.L.chol_residual(self.h, min(k, 4), C.c_uint64(seed), C.byref(r)))
        return r.value

    def residual_partial(self, k=4, seed=1):
        """this rank's share of Z = L (L^T W): (n, 4) array, permuted rows (see distributed.residual)"""
        z = np.zeros((self.n, 4), dtype=np.float64)
        self._ck(self.L.chol_residual_partial(self.h, k, C.c_uint64(seed), _p(z)))
        return z

    def residual_finish(self, z_sum, k=4, seed=1):
        z = np.ascontiguousarray(z_sum, dtype=np.float64)
        r = C.c_double()
        self._ck(self.L.chol_residual_finish(self.h, k, C.c_uint64(seed), _p(z), C.byref(r)))
        return r.value

    def matvec(self, x):
        """host check helper: A @ x from the loaded entries (original dof order)"""
        x = np.ascontiguousarray(x, dtype=np.float64).reshape(-1)
        y = np.zeros(self.n, dtype=np.float64)
        self._ck(self.L.chol_matvec(self.h, _p(x), _p(y)))
        return y

    def solve(self, b):
        b = np.ascontiguousarray(b, dtype=np.float64).reshape(-1)
        x = np.zeros(self.n, dtype=np.float64)
        self._ck(self.L.chol_solve(self.h, _p(b), _p(x)))
        return x

    # ---- solve on a partitioned handle (see cholesky_b200/distributed.py: solve)
    def solve_top_size(self):
        return int(self.L.chol_solve_top_size(self.h))

    def solve_forward(self, b):
        """forward sweep of this rank's subtree; returns its contribution to the shared top rows"""
        b = np.ascontiguousarray(b, dtype=np.float64).reshape(-1)
        top = np.zeros(max(self.solve_top_size(), 1), dtype=np.float64)
        self._ck(self.L.chol_solve_forward(self.h, _p(b), _p(top)))
        return top[:self.solve_top_size()]

    def solve_backward(self, top_sum):
        """top levels + backward sweep of the subtree; returns the entries of x this rank owns (zeros elsewhere)"""
        t = np.ascontiguousarray(top_sum, dtype=np.float64).reshape(-1)
        if t.size == 0:
            t = np.zeros(1)
        x = np.zeros(self.n, dtype=np.float64)
        self._ck(self.L.chol_solve_backward(self.h, _p(t), _p(x)))
        return x

    def solve_stats(self):
        out = np.zeros(12, dtype=np.float64)
        self._ck(self.L.chol_solve_stats(self.h, _p(out)))
        keys = ("tiles_fwd", "gemv_fwd", "pull", "gather", "tiles_bwd", "gemv_bwd")
        return dict(subtree={k: int(v) for k, v in zip(keys, out[:6])}, top={k: int(v) for k, v in zip(keys, out[6:])})


def read_vector(path, n):
    out = np.zeros(n, dtype=np.float64)
    if _lib.load().chol_read_vector(path.encode(), n, _p(out)) != 0:
        raise CholeskyError("cannot read " + path)
    return out


def write_solution(path, x):
    x = np.ascontiguousarray(x, dtype=np.float64)
    if _lib.load().chol_write_solution(path.encode(), x.size, _p(x)) != 0:
        raise CholeskyError("cannot write " + path)


def factor_binary_to_mtx(bin_path, mtx_path, full_precision=False):
    """streaming conversion of a binary factor dump to the reference's text format (host only)"""
    rc = _lib.load().chol_factor_binary_to_mtx(bin_path.encode(), mtx_path.encode(), 1 if full_precision else 0)
    if rc != 0:
        raise CholeskyError(f"factor_binary_to_mtx({bin_path}): error {rc}")


def read_factor_binary(path):
    """(n, I, J, V) of a binary factor dump: entries != 0, 0-based permuted coordinates"""
    with open(path, "rb") as f:
        raw = f.read()
    if raw[:8] != b"CHOLFAC1":
        raise CholeskyError(path + " is not a factor dump")
    n, ncols = np.frombuffer(raw, dtype=np.int32, count=2, offset=8)
    nrec, nnz = np.frombuffer(raw, dtype=np.int64, count=2, offset=24)
    off = 40
    I, J, V = [], [], []
    for _ in range(int(nrec)):
        r0, c0, nr, nc = np.frombuffer(raw, dtype=np.int32, count=4, offset=off)
        off += 16
        v = np.frombuffer(raw, dtype=np.float64, count=int(nr) * int(nc), offset=off).reshape(nc, nr).T
        off += 8 * int(nr) * int(nc)
        ii, jj = np.nonzero(v)
        I.append(ii + r0), J.append(jj + c0), V.append(v[ii, jj])
    I = np.concatenate(I) if I else np.zeros(0, dtype=np.int64)
    J = np.concatenate(J) if J else np.zeros(0, dtype=np.int64)
    V = np.concatenate(V) if V else np.zeros(0)
    if I.size != nnz:
        raise CholeskyError(path + ": entry count differs from the header")
    return int(n), I, J, V
