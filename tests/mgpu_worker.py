"""torchrun worker for the multi-GPU parity test: every rank factors its part of a generated grid on
its GPU, rank 0 gathers the factor entries of all ranks and compares them with the CPU oracle."""
import json
import os
import sys
import tempfile

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import compare_coo  # noqa: E402

from cholesky_b200 import Cholesky  # noqa: E402
from cholesky_b200.distributed import exchange_peers, make_partitioned, max_over_ranks, residual, solve  # noqa: E402


def main():
    grid = tuple(int(x) for x in sys.argv[1].split(","))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    ch = make_partitioned(grid=grid)
    exchange_peers(ch)
    st = ch.factor(iterations=2, warmup=1)
    secs = max_over_ranks(st.seconds_best)
    copies = max_over_ranks(ch.top_copies_diff())   # every rank against the next one, through peer memory
    res = residual(ch, k=4)                          # GPU-side ||(A - L L^T) W|| / ||A W||, summed over the ranks
    # triangular solve on the partitioned factor (every rank gets the whole x)
    rhs = np.random.default_rng(0).integers(1, 11, size=ch.n).astype(np.float64)
    x = solve(ch, rhs)
    entries = ch.n <= 200000   # beyond that: size-independent properties only
    # factor entries: every rank saves what it reports, rank 0 reads the files (an all_gather_object of 10^7 entries is slow)
    box = [tempfile.mkdtemp() if rank == 0 else None]
    dist.broadcast_object_list(box, 0)
    if entries:
        I, J, V = ch.factor_coo()
        np.savez(os.path.join(box[0], f"part{rank}.npz"), I=I, J=J, V=V)
    parts = [None] * world
    dist.all_gather_object(parts, ch.partition_stats())
    dist.barrier()
    ok, msg = True, ""
    if rank == 0:
        from oracle import oracle as orc
        tmp = tempfile.mkdtemp()
        m, o, c = (os.path.join(tmp, x) for x in ("a.mtx", "a_ord.txt", "a_clust.txt"))
        Cholesky(local).generate(*grid).write_inputs(m, o, c)
        ref = orc.Oracle(m, o, c)
        ref.factor(threads=4)
        if entries:
            Io, Jo, Vo = ref.factor_coo()
            zs = [np.load(os.path.join(box[0], f"part{r}.npz")) for r in range(world)]
            got = tuple(np.concatenate([z[k] for z in zs]) for k in "IJV")   # every entry is reported by exactly one rank
            ok, worst = compare_coo(ch.n, got, (Io, Jo, Vo))
            msg = f"pattern equal {ok}"
            if ok:
                ok = worst <= 1e-10
                msg = f"worst entry error {worst:.3e}"
        if ok:
            xo = ref.solve(rhs)
            sworst = float(np.max(np.abs(x - xo)) / np.max(np.abs(xo)))
            sres = float(np.linalg.norm(rhs - ch.matvec(x)) / np.linalg.norm(rhs))
            ok = sworst <= 1e-10 and sres <= 1e-12 and res <= 1e-12 and copies == 0.0
            msg += f"; solve vs oracle {sworst:.3e}, solve residual {sres:.3e}, factor residual {res:.3e}, top copies differ by {copies:.1e}"
        print(json.dumps({"ok": bool(ok), "msg": msg, "world": world, "grid": grid, "seconds": secs,
                          "gflops": ch.flops() / secs * 1e-9, "push_launches": [p["push_launches"] for p in parts], "residual": res, "copies_diff": copies}))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
