#!/usr/bin/env python
"""Join an ncu launch list (--metrics gpu__time_duration.sum --csv) with the engine's compiled launch
table and write a per-(kernel, tree level, phase) summary: launches, device time, share, TFLOP/s.

  python tools/summarize_launches.py gpurun_out/launches_64.csv lapl3d_7pt_64 > profiles/launches_64_r01.md
"""
import collections
import csv
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS  # noqa: E402
from cholesky_b200 import Cholesky  # noqa: E402

OURS = ("gemm_grouped", "gemm_small_warp", "panel_kernel", "trsm_tile")  # kernels of the launch table
PHASE = {1: "fused_dpotrf", 2: "fused_dtrsm", 3: "fused_dpotrf+dtrsm", 4: "fused_dsyrk/dgemm"}


def main():
    path, workload = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    scale = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}
    ours = [(r[ki], float(r[vi].replace(",", "")) * scale[r[ui]]) for r in rows[1:]
            if any(k in r[ki] for k in OURS)]
    other = [(r[ki], float(r[vi].replace(",", "")) * scale[r[ui]]) for r in rows[1:]
             if not any(k in r[ki] for k in OURS)]
    ch = Cholesky().generate(*WORKLOADS[workload]).analyze()
    ls = [l for l in ch.launches() if l["kind"] in ("gemm_grouped", "panel_kernel", "trsm_tile")]
    n = min(len(ls), len(ours))
    agg = collections.OrderedDict()
    tot = 0.0
    for l, (name, ms) in zip(ls[:n], ours[:n]):
        want = "gemm_small_warp" if (l["kind"] == "gemm_grouped" and l["cfg"] == 3) else l["kind"]
        assert want in name, (l, name)
        key = (want, l["level"], PHASE[l["phase"]])
        a = agg.setdefault(key, [0, 0.0, 0.0, 0])
        a[0] += 1
        a[1] += ms
        a[2] += l["flops"]
        a[3] += l["ctas"]
        tot += ms
    print(f"# ncu launch list, {workload}: {n} launches of one factorization, {tot:.3f} ms of kernel time "
          f"(cold-cache, serialised by ncu: compare shares, not absolutes)\n")
    print(f"algorithmic flops {ch.flops():.4g}; other kernels in the capture: "
          + ", ".join(f"{k.split('(')[0]} {ms:.3f} ms" for k, ms in other[:4]) + "\n")
    bykind = collections.defaultdict(float)
    print("| kernel | tree level | reference task | launches | CTAs | ms | share | TFLOP/s |")
    print("|---|---|---|---|---|---|---|---|")
    for (kind, lvl, ph), (cnt, ms, fl, ctas) in agg.items():
        bykind[kind.split("/")[0]] += ms
        tf = f"{fl / (ms * 1e-3) * 1e-12:.2f}" if fl > 0 and ms > 0 else "-"
        print(f"| {kind} | {lvl} | {ph} | {cnt} | {ctas} | {ms:.3f} | {100 * ms / tot:.1f}% | {tf} |")
    print("\n| kernel | ms | share of step |")
    print("|---|---|---|")
    for k, ms in bykind.items():
        print(f"| {k} | {ms:.3f} | {100 * ms / tot:.1f}% |")


if __name__ == "__main__":
    main()
