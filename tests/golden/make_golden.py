"""Generate tests/golden/fixtures.npz -- run in the BUILD container only (reads /root/reference).

What goes in, per reference fixture (tests/lapl_*):
  * the four input files (matrix, ord, clust, rhs) as raw bytes, so the tests can materialise
    them on a box where /root/reference does not exist;
  * golden outputs computed by the REFERENCE'S OWN checker, imported unmodified:
      verify.permute_matrix(mtx, ord)           -> permuted lower-triangular matrix (COO)
      scipy.linalg.cholesky(pmat, lower=True)   -> the factor verify.check_matrix compares with (COO)
      scipy.linalg.solve(A, b)                  -> the solution verify.check_solution compares with
  * nothing produced by this repo's oracle or CUDA path.
Usage: python tests/golden/make_golden.py
"""
import os
import sys
import warnings

import numpy as np
import scipy.io
import scipy.linalg

REF = "/root/reference"
sys.path.insert(0, REF)
import verify  # noqa: E402  (the reference's verify.py, unmodified)

CASES = {
    "lapl_9x9": ("lapl_3_2.mtx", "lapl_3_2_ord_2.txt", "lapl_3_2_clust_2.txt", "B_9x1.mtx"),
    "lapl_25x25": ("lapl_5_2.mtx", "lapl_5_2_ord_3.txt", "lapl_5_2_clust_3.txt", "B_25x1.mtx"),
    "lapl_400x400": ("lapl_20_2.mtx", "lapl_20_2_ord_5.txt", "lapl_20_2_clust_5.txt", "B_400x1.mtx"),
    "lapl_3375x3375": ("lapl_15_3.mtx", "lapl_15_3_ord_5.txt", "lapl_15_3_clust_5.txt", "B_3375x1.mtx"),
}


def main():
    warnings.simplefilter("ignore")
    out = {}
    for case, (mtx, ordf, clust, b) in CASES.items():
        d = os.path.join(REF, "tests", case)
        for kind, name in zip(("mtx", "ord", "clust", "b"), (mtx, ordf, clust, b)):
            with open(os.path.join(d, name), "rb") as f:
                out[f"{case}/file/{kind}"] = np.frombuffer(f.read(), dtype=np.uint8)
            out[f"{case}/name/{kind}"] = np.array(name)
        _, pmat = verify.permute_matrix(os.path.join(d, mtx), os.path.join(d, ordf))
        L = scipy.linalg.cholesky(pmat, lower=True)
        pi, pj = np.nonzero(pmat)
        out[f"{case}/pmat/I"], out[f"{case}/pmat/J"], out[f"{case}/pmat/V"] = pi.astype(np.int32), pj.astype(np.int32), pmat[pi, pj]
        # structural nonzeros of the dense factor: exact zeros stay exact in LAPACK's blocked potrf only
        # by luck, so threshold far below any true entry of these Laplacians (smallest |L_ij| ~ 1e-9)
        li, lj = np.nonzero(np.abs(L) > 1e-13)
        out[f"{case}/L/I"], out[f"{case}/L/J"], out[f"{case}/L/V"] = li.astype(np.int32), lj.astype(np.int32), L[li, lj]
        A = scipy.io.mmread(os.path.join(d, mtx)).toarray()
        bv = np.asarray(scipy.io.mmread(os.path.join(d, b)), dtype=np.float64).reshape(-1)
        out[f"{case}/x"] = scipy.linalg.solve(A, bv)
        print(case, "n", pmat.shape[0], "nnz(pmat)", pi.size, "nnz(L)", li.size)
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "fixtures.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
