import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CASES = ["lapl_9x9", "lapl_25x25", "lapl_400x400", "lapl_3375x3375"]
# golden structure numbers (re-derived by the oracle, pinned by tests/test_oracle.py)
GOLDEN_STRUCT = {
    "lapl_9x9": dict(n=9, nz=21, levels=2, nsep=3, blocks=5, clusters0=5, filled=[5, 5], nnzL=28,
                     calls=dict(potrf=3, trsm=2, syrk=2, gemm=0)),
    "lapl_25x25": dict(n=25, nz=65, levels=3, nsep=7, blocks=17, clusters0=37, filled=[27, 28, 1], nnzL=117,
                       calls=dict(potrf=7, trsm=16, syrk=16, gemm=14)),
    "lapl_400x400": dict(n=400, nz=1160, levels=5, nsep=31, blocks=129, clusters0=647,
                         filled=[225, 285, 69, 14, 1], nnzL=5069,
                         calls=dict(potrf=31, trsm=139, syrk=139, gemm=304)),
    "lapl_3375x3375": dict(n=3375, nz=12825, levels=5, nsep=31, blocks=129, clusters0=4776,
                           filled=[1405, 1846, 380, 44, 1], nnzL=353683,
                           calls=dict(potrf=31, trsm=425, syrk=425, gemm=3252)),
}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


class Fixture:
    """one reference fixture materialised from tests/golden/fixtures.npz into a temp dir"""

    def __init__(self, case, z, tmp):
        self.case = case
        self.dir = os.path.join(tmp, case)
        os.makedirs(self.dir, exist_ok=True)
        self.paths = {}
        for kind in ("mtx", "ord", "clust", "b"):
            name = str(z[f"{case}/name/{kind}"])
            p = os.path.join(self.dir, name)
            with open(p, "wb") as f:
                f.write(z[f"{case}/file/{kind}"].tobytes())
            self.paths[kind] = p
        self.mtx, self.ord, self.clust, self.b = (self.paths[k] for k in ("mtx", "ord", "clust", "b"))
        self.pmat = (z[f"{case}/pmat/I"], z[f"{case}/pmat/J"], z[f"{case}/pmat/V"])
        self.L = (z[f"{case}/L/I"], z[f"{case}/L/J"], z[f"{case}/L/V"])
        self.x = z[f"{case}/x"]
        self.struct = GOLDEN_STRUCT[case]
        self.n = self.struct["n"]

    def L_dense(self):
        d = np.zeros((self.n, self.n))
        d[self.L[0], self.L[1]] = self.L[2]
        return d

    def pmat_dense(self):
        d = np.zeros((self.n, self.n))
        d[self.pmat[0], self.pmat[1]] = self.pmat[2]
        return d


@pytest.fixture(scope="session")
def golden(tmp_path_factory):
    z = np.load(os.path.join(ROOT, "tests", "golden", "fixtures.npz"))
    tmp = str(tmp_path_factory.mktemp("fixtures"))
    return {c: Fixture(c, z, tmp) for c in CASES}


def entrywise_ok(L, Lref, rtol=1e-10, floor=1e-6):
    """the parity rule: |dL_ij| <= rtol * max(|Lref_ij|, floor * max|Lref|)"""
    scale = np.maximum(np.abs(Lref), floor * np.abs(Lref).max())
    return float(np.max(np.abs(L - Lref) / scale)) <= rtol, float(np.max(np.abs(L - Lref) / scale))


def compare_coo(n, got, want, rtol_floor=1e-6):
    """(pattern identical, worst entry error by the parity rule) of two COO factors, vectorised (10^8 entries are fine)"""
    def canon(t):
        I, J, V = t
        key = I.astype(np.int64) * n + J.astype(np.int64)
        o = np.argsort(key, kind="stable")
        return key[o], np.asarray(V)[o]
    kg, vg = canon(got)
    kw, vw = canon(want)
    if kg.shape != kw.shape or not np.array_equal(kg, kw):
        return False, float("inf")
    scale = np.maximum(np.abs(vw), rtol_floor * np.abs(vw).max())
    return True, float(np.max(np.abs(vg - vw) / scale))
