#!/usr/bin/env python
"""Parity + timing of the experimental kernel variants (default off) in one process, a few seconds of GPU:
for every variant, factor the 15^3 fixture and compare with the golden factor, then time 64^3 and 512^2.
  python tools/check_experimental.py            (needs a GPU)"""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cholesky_b200 import Cholesky  # noqa: E402

VARIANTS = [{}, {"CHOL_POTRF_R": "2"}, {"CHOL_POTRF_R": "3"}, {"CHOL_POTRF_R": "1"}, {"CHOL_TRSM_BATCH": "1"}, {"CHOL_POTRF_R": "2", "CHOL_TRSM_BATCH": "1"},
            {"CHOL_GEMM_STAGES": "4"}, {"CHOL_NBO_SMALL": "128"}, {"CHOL_GRAPH": "1"},
            {"CHOL_GRAPH": "1", "CHOL_POTRF_R": "2", "CHOL_TRSM_BATCH": "1"}]
KEYS = ("CHOL_POTRF_R", "CHOL_TRSM_BATCH", "CHOL_GEMM_STAGES", "CHOL_NBO_SMALL", "CHOL_GRAPH")


def main():
    z = np.load(os.path.join(ROOT, "tests", "golden", "fixtures.npz"))
    case = "lapl_3375x3375"
    d = tempfile.mkdtemp()
    paths = {}
    for kind in ("mtx", "ord", "clust"):
        paths[kind] = os.path.join(d, str(z[f"{case}/name/{kind}"]))
        with open(paths[kind], "wb") as f:
            f.write(z[f"{case}/file/{kind}"].tobytes())
    n = 3375
    Lg = np.zeros((n, n))
    Lg[z[f"{case}/L/I"], z[f"{case}/L/J"]] = z[f"{case}/L/V"]
    scale = np.maximum(np.abs(Lg), 1e-6 * np.abs(Lg).max())
    for v in VARIANTS:
        for k in KEYS:
            os.environ.pop(k, None)
        os.environ.update(v)
        out = {"variant": v}
        try:
            ch = Cholesky().load(paths["mtx"], paths["ord"], paths["clust"]).analyze()
            ch.factor()
            out["max_entry_error"] = float(np.max(np.abs(ch.factor_dense() - Lg) / scale))
            ch.close()
            big = Cholesky().generate(64, 64, 64, 7, 0).analyze()
            st = big.factor(iterations=3, warmup=1)
            out["ms_64"] = st.seconds_best * 1e3
            out.update({k: round(x, 3) for k, x in big.kernel_times().items() if k.endswith("_ms")})
            big.close()
            flat = Cholesky().generate(512, 512, 1, 5, 0).analyze()   # the launch-bound BASELINE config 2
            out["ms_512sq"] = flat.factor(iterations=5, warmup=2).seconds_best * 1e3
            flat.close()
        except Exception as e:  # noqa: BLE001
            out["error"] = str(e)
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
