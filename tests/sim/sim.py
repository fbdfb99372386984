"""ctypes wrapper of tests/sim/libsched_sim.so: the host interpreter of the engine's launch lists
(TEST INFRASTRUCTURE ONLY; see sched_sim.cc)."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libsched_sim.so")


def _lib():
    subprocess.check_call(["make", "-C", HERE, "-s"])
    return C.CDLL(LIB)


def factor(grid, world, seed=1):
    """(dense L in permuted order, max |difference| between the ranks' copies of the top panels, stats)"""
    nx, ny, nz, stencil, levels = grid
    n = nx * ny * nz
    L = np.zeros((n, n))
    diff = C.c_double()
    stats = np.zeros(4)
    err = C.create_string_buffer(4096)
    rc = _lib().sim_factor(nx, ny, nz, stencil, levels, world, C.c_uint64(seed), L.ctypes.data_as(C.c_void_p), C.byref(diff),
                           stats.ctypes.data_as(C.c_void_p), err, 4096)
    if rc != 0:
        raise RuntimeError("sched_sim: " + err.value.decode())
    return L, diff.value, dict(push_rects=int(stats[0]), syncs=int(stats[1]), reduces=int(stats[2]), gemm_flops=float(stats[3]))
