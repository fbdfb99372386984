#!/usr/bin/env python
"""One factorization of a named workload, for ncu / launch-list captures.

  python tools/profile_step.py --workload lapl3d_7pt_64 [--list] [--iterations 1] [--warmup 0]

--list only analyses (no GPU needed) and prints the launch table summary plus, for every kernel, the
ordinal (among launches of that kernel) of its three largest launches -- the numbers to give ncu's
`-k regex:<kernel> -s <ordinal> -c 1`.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS  # noqa: E402
from cholesky_b200 import Cholesky  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="lapl3d_7pt_64")
    ap.add_argument("--list", action="store_true")
    ap.add_argument("--iterations", type=int, default=1)
    ap.add_argument("--warmup", type=int, default=0)
    a = ap.parse_args()
    ch = Cholesky(0).generate(*WORKLOADS[a.workload]).analyze()
    ls = ch.launches()
    by = {}
    for l in ls:
        by.setdefault(l["kind"], []).append(l)
    summary = {"workload": a.workload, "launches": len(ls), "flops": ch.flops()}
    for k, v in by.items():
        top = sorted(range(len(v)), key=lambda i: -(v[i]["flops"] if k == "gemm_grouped" else v[i]["ctas"]))[:3]
        summary[k] = {"count": len(v), "ctas": sum(x["ctas"] for x in v),
                      "top": [{"ordinal": i, **v[i]} for i in top]}
    print(json.dumps(summary))
    if a.list:
        return
    st = ch.factor(iterations=a.iterations, warmup=a.warmup)
    print(json.dumps({"seconds_best": st.seconds_best, "gflops": st.flops / st.seconds_best * 1e-9,
                      "kernel_launches": st.kernel_launches, "info": st.info}))


if __name__ == "__main__":
    main()
