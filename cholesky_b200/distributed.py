"""Multi-GPU plumbing for the one-process-per-GPU form (torchrun); torch.distributed only for the bootstrap.
(Inside one process, Cholesky(devices=[...]) drives the same partition through a group handle.)

Rank r owns the subtree under heap index world + r; the rows of the log2(world) top levels' panels are dealt
to the ranks under each separator in blocks of 256 (csrc/schedule.cc).  The data path never goes
through torch or NCCL: after the ranks have exchanged the CUDA-IPC handles of their factor buffers and flag words
(one all_gather of 128 bytes per rank), the engine's own kernels pull the partial sums of the rows a rank owns,
push factored rows into the peers' copies and meet in flag words, all through NVLink peer memory (csrc/kernels.cuh).
"""
import os
import tempfile

import torch
import torch.distributed as dist

from .engine import Cholesky


def make_partitioned(grid=None, files=None, keep_records=False):
    """build this rank's engine: generate/load, set the partition, analyse.  The symbolic analysis is the same on
    every rank, so rank 0 runs it with all the host threads and hands the result to the others through a file in
    shared memory (CHOL_SHARE_ANALYSIS=0: every rank analyses for itself, side by side on a share of the cores)"""
    rank, world = dist.get_rank(), dist.get_world_size()
    share = world > 1 and os.environ.get("CHOL_SHARE_ANALYSIS", "1") != "0"
    if not share:  # the ranks of one node share its host cores
        os.environ.setdefault("CHOL_HOST_THREADS", str(max(1, min(16, (os.cpu_count() or 1) // world))))
    dev = torch.cuda.current_device() if torch.cuda.is_available() else 0
    ch = Cholesky(dev)
    if grid is not None:
        ch.generate(*grid)
    else:
        ch.load(*files)
    ch.set_partition(rank, world)
    if not share:
        return ch.analyze(keep_records=keep_records)
    box = [None]
    if rank == 0:
        ch.analyze(keep_records=keep_records)
        for d in ("/dev/shm", tempfile.gettempdir()):   # (15 MB at 128^3)
            path = os.path.join(d, f"chol_analysis_{os.getpid()}_{os.environ.get('MASTER_PORT', '0')}.bin")
            try:
                ch.save_analysis(path)
                box[0] = path
                break
            except Exception:  # noqa: BLE001 -- no room or no such directory: try the next place, else every rank analyses
                pass
    dist.broadcast_object_list(box, 0)
    if rank != 0:
        try:
            if box[0] is None:
                raise FileNotFoundError
            ch.load_analysis(box[0])
        except Exception:  # noqa: BLE001
            ch.analyze(keep_records=keep_records)
    dist.barrier()
    if rank == 0 and box[0] is not None:
        os.unlink(box[0])
    return ch


def exchange_peers(ch):
    """all-gather the IPC handles and map every peer's buffers into this rank"""
    world = dist.get_world_size()
    if world == 1:
        return
    blob = ch.ipc_export()
    backend = dist.get_backend()
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    mine = torch.tensor(list(blob), dtype=torch.uint8, device=device)
    allb = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allb, mine)
    ch.ipc_import([bytes(t.cpu().tolist()) for t in allb])
    dist.barrier()


def max_over_ranks(x):
    if dist.is_initialized() and dist.get_world_size() > 1:
        device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.tensor([x], dtype=torch.float64, device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())
    return x


def _sum_over_ranks(a):
    """element-wise sum of a float64 numpy array over the ranks (bootstrap-grade: staged through torch)"""
    import numpy as np
    if not dist.is_initialized() or dist.get_world_size() == 1 or a.size == 0:
        return a
    device = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    t = torch.from_numpy(np.ascontiguousarray(a)).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()


def solve(ch, b):
    """A x = b on a partitioned factor (mmat.rg:1364-1495 over the ranks' subtrees): every rank passes the
    same b and gets the whole x.  The sweeps run on each rank's GPU; two small host-staged sums cross the
    ranks -- the top part of the right-hand side after the subtree forward sweeps (chol_solve_top_size
    doubles) and the assembly of the owned pieces of x at the end."""
    top = _sum_over_ranks(ch.solve_forward(b))
    return _sum_over_ranks(ch.solve_backward(top))


def residual(ch, k=4, seed=1):
    """randomized ||(A - L L^T) W||_F / ||A W||_F of a partitioned factor: every rank evaluates its panels' share of
    L (L^T W) on its GPU, the (n, 4) shares are summed over the ranks, the comparison with A W runs on the host"""
    z = _sum_over_ranks(ch.residual_partial(k=k, seed=seed))
    return ch.residual_finish(z, k=k, seed=seed)
