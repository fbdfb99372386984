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
from cholesky_b200 import Cholesky  # noqa: E402
from cholesky_b200.distributed import exchange_peers, make_partitioned, max_over_ranks, solve  # noqa: E402


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
    # triangular solve on the partitioned factor (every rank gets the whole x)
    rhs = np.random.default_rng(0).integers(1, 11, size=ch.n).astype(np.float64)
    x = solve(ch, rhs)
    I, J, V = ch.factor_coo()
    parts = [None] * world
    dist.all_gather_object(parts, (I, J, V, ch.partition_stats()))
    ok, msg = True, ""
    if rank == 0:
        from oracle import oracle as orc
        tmp = tempfile.mkdtemp()
        m, o, c = (os.path.join(tmp, x) for x in ("a.mtx", "a_ord.txt", "a_clust.txt"))
        Cholesky(local).generate(*grid).write_inputs(m, o, c)
        ref = orc.Oracle(m, o, c)
        ref.factor(threads=4)
        Io, Jo, Vo = ref.factor_coo()
        want = {(int(i), int(j)): float(v) for i, j, v in zip(Io, Jo, Vo)}
        got = {}
        for (pi, pj, pv, _) in parts:
            for i, j, v in zip(pi.tolist(), pj.tolist(), pv.tolist()):
                assert (i, j) not in got, "entry reported by two ranks"
                got[(i, j)] = v
        ok = got.keys() == want.keys()
        msg = f"keys equal {ok}"
        if ok:
            a = np.array([got[k] for k in want])
            b = np.array([want[k] for k in want])
            worst = float(np.max(np.abs(a - b) / np.maximum(np.abs(b), 1e-6 * np.abs(b).max())))
            ok = worst <= 1e-10
            msg = f"worst entry error {worst:.3e}"
        if ok:
            xo = ref.solve(rhs)
            sworst = float(np.max(np.abs(x - xo)) / np.max(np.abs(xo)))
            res = float(np.linalg.norm(rhs - ch.matvec(x)) / np.linalg.norm(rhs))
            ok = sworst <= 1e-10 and res <= 1e-12
            msg += f"; solve vs oracle {sworst:.3e}, residual {res:.3e}"
        print(json.dumps({"ok": bool(ok), "msg": msg, "world": world, "grid": grid, "seconds": secs,
                          "gflops": ch.flops() / secs * 1e-9, "shared_launches": [p[3]["shared_launches"] for p in parts]}))
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()
