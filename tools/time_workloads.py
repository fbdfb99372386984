#!/usr/bin/env python
"""Step time and per-kernel-class device times of the BASELINE workloads on one GPU (one JSON line each).
  python tools/time_workloads.py [workload ...]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS  # noqa: E402
from cholesky_b200 import Cholesky  # noqa: E402

for name in (sys.argv[1:] or ["lapl2d_5pt_512", "lapl3d_7pt_64", "lapl3d_27pt_96", "lapl3d_7pt_128"]):
    ch = Cholesky(0).generate(*WORKLOADS[name]).analyze()
    st = ch.factor(iterations=5 if ch.n < 1000000 else 3, warmup=2)
    out = {"workload": name, "ms": round(st.seconds_best * 1e3, 3), "tflops": round(st.flops / st.seconds_best * 1e-12, 2),
           "launches": int(st.kernel_launches)}
    out.update({k: round(v, 3) for k, v in ch.kernel_times().items() if k.endswith("_ms")})
    out["residual"] = ch.residual(k=2)
    print(json.dumps(out), flush=True)
    ch.close()
