#!/usr/bin/env python
"""bench.py -- factor GFLOP/s & time (FP64, 3-D Laplacian) of the numeric sparse Cholesky path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload lapl3d_7pt_128] [--impl reference]

One "step" = one numeric factorization (assemble + level loop) of the workload.  `value` is algorithmic
GFLOP/s (flops of the reference's BLAS call list / device time of the level loop, CUDA events on the
launching stream, inputs resident in HBM); `e2e` is the same metric through the C-ABI call that takes
HOST buffers (H2D of A's values, assemble, factor, D2H of diag(L) inside the timed region).
`--impl reference` times the reference's algorithm on the host cores (CPU oracle over host OpenBLAS,
the only place besides cpu_baseline where oracle/ is executed): ONE factorization of the full workload
when the host has the cores and memory for it (128^3: ~26 GB, about a minute of BLAS on 16 cores),
otherwise a smaller grid -- `config.workload` always names the grid that was actually factored.
With more than one rank the line also carries the correctness of that very factor: GPU-side randomized
residual, partitioned solve residual, and the largest difference between the ranks' copies of the top panels.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (nx, ny, nz, stencil, levels[0 = utils.py rule])
    "lapl3d_7pt_128": (128, 128, 128, 7, 0),   # BASELINE config 4 (the headline)
    "lapl3d_7pt_64": (64, 64, 64, 7, 0),       # config 3
    "lapl2d_5pt_512": (512, 512, 1, 5, 0),     # config 2
    "lapl3d_27pt_96": (96, 96, 96, 27, 0),     # config 5
    "lapl3d_7pt_15": (15, 15, 15, 7, 5),       # config 1 grid (generated ordering)
}
# bounded CPU sample per workload: (grid, description)
CPU_SAMPLE = {
    "lapl3d_7pt_128": ((80, 80, 80, 7, 0), "80^3 7-pt Laplacian (a quarter of the 128^3 workload's unknowns, 1/17 of its "
                                           "flops), full factorization, same generator"),
    "lapl3d_7pt_64": ((48, 48, 48, 7, 0), "48^3 7-pt Laplacian, full factorization, same generator"),
    "lapl2d_5pt_512": ((512, 512, 1, 5, 0), "the full 512x512 5-pt workload"),
    "lapl3d_27pt_96": ((40, 40, 40, 27, 0), "40^3 27-pt Laplacian, full factorization, same generator"),
    "lapl3d_7pt_15": ((15, 15, 15, 7, 5), "the full 15^3 workload"),
}
METRIC = "factor GFLOP/s (FP64, 3D Laplacian numeric sparse Cholesky)"


def fp64_peak():
    """FP64 peak in TFLOP/s.  MEASURED_PEAKS.json carries no FP64 figure, so the denominator is this
    repo's own measurement on the pool's B200 (tools/fp64_peak.cu, summary in profiles/)."""
    p = os.path.join(ROOT, "profiles", "fp64_peak.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["fp64_tflops"]), d.get("how", "profiles/fp64_peak.json")
    return 37.0, "nominal HGX B200 FP64 (296 TF / 8 GPUs); no measurement committed yet"


class ClockSampler:
    def __init__(self):
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
                 "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap",
                 "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self, device=0):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        rows = [r for r in self.rows if len(r) >= 9 and r[0] == str(device)]
        if not rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        sm = sorted(float(r[1]) for r in rows)
        reasons = set()
        for r in rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(rows[0][2]), "reasons": sorted(reasons),
                "samples": len(rows), "power_w_max": max(float(r[3]) for r in rows)}


def describe(name, grid):
    return (f"{name}: {grid[0]}x{grid[1]}x{grid[2]} grid, {grid[3]}-point Laplacian, geometric ND "
            f"(levels by utils.py rule), reference-format ord/clust generated in memory")


def host_can_run_full(workload):
    """the oracle keeps the whole factor in host memory (the reference's filled clusters, 128^3: 25.5 GiB) and needs
    about a minute of multi-threaded BLAS for 128^3 on 16 cores"""
    if os.environ.get("CHOL_REF_SAMPLE") in ("0", "1"):    # 1: always the bounded sample, 0: always the full workload
        return os.environ["CHOL_REF_SAMPLE"] == "0"
    need_gb = {"lapl3d_7pt_128": 48, "lapl3d_27pt_96": 24}.get(workload, 8)
    try:
        avail = [int(l.split()[1]) for l in open("/proc/meminfo") if l.startswith("MemAvailable")][0] / 2**20
    except Exception:
        avail = 0
    return avail >= need_gb and (os.cpu_count() or 1) >= 12


def cpu_reference(workload, threads=None, repeats=1, full=False):
    """the reference's blocked algorithm over host BLAS (oracle/), timed on this box's host cores"""
    from cholesky_b200 import Cholesky
    from oracle import oracle as orc
    grid, desc = CPU_SAMPLE[workload]
    if full:
        grid, desc = WORKLOADS[workload], "the full workload, one factorization"
    threads = threads or os.cpu_count() or 1
    tmp = tempfile.mkdtemp()
    m, o, c = (os.path.join(tmp, x) for x in ("a.mtx", "a_ord.txt", "a_clust.txt"))
    Cholesky().generate(*grid).write_inputs(m, o, c)   # input generation only; no numeric call
    ref = orc.Oracle(m, o, c)
    best = None
    for _ in range(repeats):
        s = ref.factor(threads=threads)
        best = s if best is None else min(best, s)
    return {"value": ref.flops() / best * 1e-9, "unit": "GFLOP/s", "cores": threads, "kind": "port",
            "sample": desc + f"; {ref.flops():.4g} flops in {best:.2f} s; host BLAS {orc.blas_config()}",
            "seconds": best, "grid": grid}


def measured_hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return float(json.load(open(p))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    return 6554.6, "round-1 MEASURED_PEAKS.json value (file absent on this box)"


def small_front_roofline(ch, first_level=8):
    """HBM view of the bottom of the tree: algorithmic bytes (chol_level_bytes: panels factored in place + distinct
    Schur operands + destinations read-modify-written) of tree levels >= first_level over their device time in the
    instrumented pass of THIS run"""
    if ch.levels <= first_level:
        return None
    ls, ms = ch.launches(), ch.launch_times()
    t = sum(float(m) for l, m in zip(ls, ms) if l["level"] >= first_level) * 1e-3
    b = sum(sum(ch.level_bytes(lv).values()) for lv in range(first_level, ch.levels))
    peak, how = measured_hbm_peak()
    if t <= 0:
        return None
    return {"bound": "hbm", "levels": f">= {first_level}", "achieved": b / t * 1e-9, "peak": peak, "unit": "GB/s",
            "frac": b / t * 1e-9 / peak, "bytes": b, "ms": t * 1e3, "peak_source": how}


def ncu_traffic(workload):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of this workload
    (profiles/ncu_traffic.json: {workload: {"traffic": bytes, "launch": "...", "source": "..."}}); None if not captured"""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        return json.load(open(p)).get(workload)
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=os.environ.get("CHOL_BENCH_WORKLOAD", "lapl3d_7pt_128"))
    ap.add_argument("--impl", default="ours")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    grid = WORKLOADS[args.workload]
    config = {"workload": describe(args.workload, grid),
              "l2_policy": "inputs larger than L2: every step re-assembles and rewrites the whole factor "
                           "(>= 1.5 GB, 126 MB L2)"}

    if args.impl == "reference":
        if rank != 0:
            return
        t0 = time.time()
        full = host_can_run_full(args.workload)
        # one factorization is one step; the sampled grid is small enough for up to three
        res = [cpu_reference(args.workload, full=full) for _ in range(1 if full else max(1, min(args.steps, 3)))]
        best = max(res, key=lambda r: r["value"])
        if not full:   # name the grid that was actually factored
            g = best["grid"]
            config["sample_of"] = config["workload"]
            config["workload"] = describe(f"bounded sample of {args.workload}", g)
        config["l2_policy"] = "host run"
        line = {"impl": "reference", "metric": METRIC, "value": best["value"], "unit": "GFLOP/s", "n_gpus": args.gpus,
                "steps": len(res), "warmup": 0, "ms_per_step": best["seconds"] * 1e3, "higher_is_better": True,
                "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {k: best[k] for k in ("value", "unit", "cores", "kind", "sample")},
                "e2e": {"value": best["value"], "unit": "GFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "wall_s": time.time() - t0}
        print(json.dumps(line))
        return

    import torch
    import torch.distributed as dist
    from cholesky_b200 import Cholesky

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the numeric path has no CPU fallback")
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)

    t0 = time.time()
    if world > 1:
        # one process per GPU: rank r owns the subtree under heap index world + r, the top log2(world)
        # levels are shared; data moves through NVLink peer memory inside the engine's own kernels
        from cholesky_b200.distributed import exchange_peers, make_partitioned
        ch = make_partitioned(grid=grid)      # rank 0 analyses, the others read its result
        analyze_s = time.time() - t0
        exchange_peers(ch)                    # device buffers (the whole factor allocation per rank) + CUDA-IPC mapping of the peers'
    else:
        ch = Cholesky(local_rank).generate(*grid)
        ch.analyze()
        analyze_s = time.time() - t0
        ch.assemble()                         # device buffers
    setup_s = time.time() - t0 - analyze_s
    flops = ch.flops()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler()
    barrier()
    if rank == 0:
        sampler.start()
    st = ch.factor(iterations=args.steps, warmup=args.warmup)   # device-timed per step with CUDA events
    barrier()
    # per-step time: the slowest rank
    step_s = st.seconds_median
    if world > 1:
        t = torch.tensor([step_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        step_s = float(t.item())
    # end to end through the host-buffer C-ABI call
    e2e_s = []
    for _ in range(max(1, args.steps)):
        _, st2 = ch.factor_host()
        e2e_s.append(st2.seconds_best)
    e2e_med = sorted(e2e_s)[len(e2e_s) // 2]
    if world > 1:
        t = torch.tensor([e2e_med], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_med = float(t.item())
    barrier()
    clocks = sampler.stop(local_rank) if rank == 0 else None
    kt = ch.kernel_times()      # one instrumented iteration: CUDA events around every launch (all ranks take part)
    launches = int(st.kernel_launches)
    import numpy as np
    rhs = np.random.default_rng(0).integers(1, 11, size=ch.n).astype(np.float64)
    copies_diff = None
    ch.factor()                  # the instrumented pass ran everything on one stream; check a factor of the timed kind
    t1 = time.time()
    if world > 1:
        from cholesky_b200.distributed import residual as dist_residual, solve as dist_solve
        t = torch.tensor([float(launches)], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        launches = int(t.item())
        res = dist_residual(ch, k=2)       # GPU-side ||(A - L L^T) W|| / ||A W||, every rank its own panels
        t = torch.tensor([ch.top_copies_diff()], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        copies_diff = float(t.item())
        xs = dist_solve(ch, rhs)           # partitioned forward/backward sweeps (mmat.rg:1364-1495)
        t1 = time.time()
        dist_solve(ch, rhs)
    else:
        res = ch.residual(k=2)
        xs = ch.solve(rhs)                 # forward/backward sweeps on the GPU, host vectors in and out
        t1 = time.time()
        ch.solve(rhs)
    solve_ms = (time.time() - t1) * 1e3
    # full-size correctness through the factor: ||b - A x|| / ||b|| (A x on the host from the input entries)
    solve_res = float(np.linalg.norm(rhs - ch.matvec(xs)) / np.linalg.norm(rhs))
    small = small_front_roofline(ch) if world == 1 else None

    if rank == 0:
        peak, peak_how = fp64_peak()
        gemm_tf = kt["gemm_flops"] / (kt["gemm_ms"] * 1e-3) * 1e-12 if kt["gemm_ms"] > 0 else 0.0
        tot_ms = kt["panel_ms"] + kt["exchange_ms"] + kt["gemm_ms"]
        traffic = ncu_traffic(args.workload) if world == 1 else None
        line = {
            "metric": METRIC, "value": flops / step_s * 1e-9, "unit": "GFLOP/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_s * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
            "e2e": {"value": flops / e2e_med * 1e-9, "unit": "GFLOP/s", "h2d_bytes_per_step": ch.nz * 8,
                    "d2h_bytes_per_step": ch.n * 8, "ms_per_step": e2e_med * 1e3,
                    "note": "H2D: the nz values of A; D2H: diag(L) only -- the factor stays in HBM, where the solve consumes it"},
            "gpu_launches": launches * args.steps,
            "clocks": clocks,
            "roofline": {"bound": "tensor", "kernel": "gemm_grouped_ws (FP64 DMMA)", "achieved": gemm_tf, "peak": peak,
                         "unit": "TFLOP/s", "frac": gemm_tf / peak,
                         "traffic": traffic["traffic"] if traffic else None, "traffic_source": traffic,
                         "peak_source": peak_how,
                         "kernel_share_of_step": kt["gemm_ms"] / tot_ms if tot_ms else None,
                         "kernel_ms": {k: kt[k] for k in ("panel_ms", "exchange_ms", "gemm_ms")},
                         "whole_step_frac": flops / step_s * 1e-12 / (peak * world)},
            "roofline_small": small,
            "factor": {"n": ch.n, "nz": ch.nz, "levels": ch.levels, "flops": flops, "factor_GiB": ch.factor_doubles() * 8 / 2**30,
                       "analyze_s": analyze_s, "device_setup_s": setup_s, "assemble_ms": st.assemble_seconds * 1e3, "seconds_best": st.seconds_best,
                       "residual": res, "solve_ms": solve_ms, "solve_rel_residual": solve_res, "top_copies_max_diff": copies_diff},
        }
        if res > 1e-12 or solve_res > 1e-10 or (copies_diff or 0.0) != 0.0:
            line["INVALID"] = "the factor failed its correctness checks"
        if world > 1:
            line["config"]["parallelism"] = (f"{world} ranks: one subtree per GPU below tree level {world.bit_length() - 1}; rows of the top "
                                             "panels dealt to the ranks under each separator in blocks of 256 (owner computes), partial sums "
                                             "pulled and factored rows pushed through NVLink peer memory by the engine's own kernels")
        if not args.no_cpu_baseline and world == 1:
            cb = cpu_reference(args.workload)
            line["cpu_baseline"] = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
