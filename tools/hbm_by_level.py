#!/usr/bin/env python
"""HBM roofline by tree level: join the per-launch device times of a launch report (tools/launch_report.py,
CUDA events around every launch) with the algorithmic bytes of each level (chol_level_bytes: SURVEY 8(d),
8 B x distinct clusters read + written, read-modify-written clusters twice).  Host only.

  python tools/hbm_by_level.py profiles/launch_report_128_r01b.md lapl3d_7pt_128 > profiles/hbm_by_level_128_r01.md
"""
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS  # noqa: E402
from cholesky_b200 import Cholesky  # noqa: E402


def main():
    report, workload = sys.argv[1], sys.argv[2]
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
    hbm = float(peaks.get("hbm_gbs", 6554.6))
    ch = Cholesky().generate(*WORKLOADS[workload]).analyze()
    ms = {}
    for line in open(report):
        m = re.match(r"\| ([\w/]+) \| (\d+) \| ([\w+]+) \| (\d+) \| (\d+) \| ([\d.]+) \|", line)
        if m:
            _, lvl, ph, _, _, t = m.groups()
            d = ms.setdefault(int(lvl), {"chain": 0.0, "update": 0.0})
            d["update" if ph == "update" else "chain"] += float(t)
    fl = ch.flops_by_level()
    tot = sum(float(sum(fl[k])) for k in fl)
    print(f"# HBM roofline by tree level, {workload} (times: {os.path.basename(report)}; peak {hbm:.0f} GB/s, MEASURED_PEAKS.json)\n")
    print("Panel phase = fused_dpotrf + fused_dtrsm of the level (panels factored in place, 16 B per stored entry);")
    print("update phase = fused_dsyrk/dgemm (8 B per distinct operand entry + 16 B per destination entry).\n")
    print("| tree level | flops share | panel GB | panel ms | GB/s | % of HBM peak | update GB (operands + destinations) | update ms | GB/s | % of HBM peak |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for lvl in range(ch.levels - 1, -1, -1):
        b = ch.level_bytes(lvl)
        t = ms.get(lvl, {"chain": 0.0, "update": 0.0})
        f = sum(float(fl[k][lvl]) for k in fl)
        ub = b["operands"] + b["destinations"]
        pg = b["panel"] / 1e9 / (t["chain"] * 1e-3) if t["chain"] else 0.0
        ug = ub / 1e9 / (t["update"] * 1e-3) if t["update"] else 0.0
        print(f"| {lvl} | {100 * f / tot:.2f} % | {b['panel'] / 1e9:.3f} | {t['chain']:.3f} | {pg:.0f} | {100 * pg / hbm:.1f} % | "
              f"{ub / 1e9:.3f} ({b['operands'] / 1e9:.3f} + {b['destinations'] / 1e9:.3f}) | {t['update']:.3f} | {ug:.0f} | {100 * ug / hbm:.1f} % |")


if __name__ == "__main__":
    main()
