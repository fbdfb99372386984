#!/usr/bin/env python
"""Per-launch device times of one instrumented factorization (CUDA events around every launch):
top launches by time and a per-(kernel, level, phase) table.  Needs a GPU; under torchrun it reports
rank 0's stream of the partitioned factorization.
  python tools/launch_report.py --workload lapl3d_7pt_128 > gpurun_out/launch_report_128.md"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS  # noqa: E402
from cholesky_b200 import Cholesky  # noqa: E402

PHASE = {1: "potrf", 2: "trsm", 3: "potrf+trsm", 4: "update"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="lapl3d_7pt_128")
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        import torch
        import torch.distributed as dist
        from cholesky_b200.distributed import exchange_peers, make_partitioned
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        ch = make_partitioned(grid=WORKLOADS[a.workload])
        exchange_peers(ch)
    else:
        ch = Cholesky(0).generate(*WORKLOADS[a.workload]).analyze()
    st = ch.factor(iterations=1, warmup=1)
    kt = ch.kernel_times()
    ms = ch.launch_times()
    ls = ch.launches()
    if rank == 0:
        tot = float(ms.sum())
        print(f"# per-launch CUDA-event times, {a.workload}, world {world} (rank 0): {len(ls)} launches, {tot:.2f} ms "
              f"(uninstrumented step {st.seconds_best * 1e3:.2f} ms); kernel_ms {kt}\n")
        agg = collections.OrderedDict()
        bykind = collections.defaultdict(float)
        for l, t in zip(ls, ms):
            k = (l["kind"] + ("" if l["stream"] == 0 else "/chain" if l["stream"] == 1 else "/background"), l["level"],
                 PHASE.get(l["phase"], str(l["phase"])))
            v = agg.setdefault(k, [0, 0.0, 0.0, 0])
            v[0] += 1
            v[1] += float(t)
            v[2] += l["flops"]
            v[3] += l["ctas"]
            bykind[k[0]] += float(t)
        print("| kernel | ms | share |\n|---|---|---|")
        for k, t in sorted(bykind.items(), key=lambda x: -x[1]):
            print(f"| {k} | {t:.3f} | {100 * t / tot:.1f}% |")
        print("\n| kernel | level | phase | launches | CTAs | ms | share | TFLOP/s |\n|---|---|---|---|---|---|---|---|")
        for (kind, lvl, ph), (n, t, f, ctas) in agg.items():
            tf = f"{f / (t * 1e-3) * 1e-12:.2f}" if f > 0 and t > 0 else "-"
            print(f"| {kind} | {lvl} | {ph} | {n} | {ctas} | {t:.3f} | {100 * t / tot:.1f}% | {tf} |")
        print("\n## top 25 launches\n\n| # | kernel | level | phase | CTAs | ms | TFLOP/s |\n|---|---|---|---|---|---|---|")
        order = sorted(range(len(ls)), key=lambda i: -ms[i])[:25]
        for i in order:
            l = ls[i]
            tf = f"{l['flops'] / (ms[i] * 1e-3) * 1e-12:.2f}" if l["flops"] > 0 else "-"
            print(f"| {i} | {l['kind']} | {l['level']} | {PHASE.get(l['phase'])} | {l['ctas']} | {ms[i]:.3f} | {tf} |")
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
