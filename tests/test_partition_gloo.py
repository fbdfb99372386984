"""CPU tests of the multi-GPU partition (world_size 2 and 4 over the gloo backend): every rank
compiles its own schedule; together they must cover the single-rank schedule exactly once where
work is split and identically where it is replicated.  No GPU, no compute calls."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

GRID = (40, 24, 20, 7, 6)   # 480-dof root: its panel has trailing updates (block columns of 256)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    os.environ["CHOL_SHARED_MIN_FLOPS"] = "1"   # split every top-level GEMM launch of this small grid
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from cholesky_b200 import Cholesky
        from cholesky_b200.distributed import make_partitioned, max_over_ranks
        ch = make_partitioned(grid=GRID)
        mine = dict(rank=rank, stats=ch.partition_stats(), launches=ch.launches(), flops=ch.flops(), solve=ch.solve_stats(),
                    top_size=ch.solve_top_size(),
                    checksums=[ch.filled_checksum(t) for t in range(ch.levels)], nz=ch.nz, levels=ch.levels)
        gathered = [None] * world
        dist.all_gather_object(gathered, mine)
        assert max_over_ranks(float(rank)) == world - 1
        if rank == 0:
            single = Cholesky().generate(*GRID).analyze()
            q.put(dict(ranks=gathered, single=dict(launches=single.launches(), stats=single.partition_stats(),
                                                   solve=single.solve_stats(), sizes=single.sep_sizes().tolist())))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 4])
def test_partition_covers_single_rank_schedule(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ranks, single = out["ranks"], out["single"]
    depth = world.bit_length() - 1
    # every rank analysed the same pattern
    assert len({tuple(r["checksums"]) for r in ranks}) == 1
    # each matrix entry is assembled by exactly one rank
    assert sum(r["stats"]["assembled"] for r in ranks) == ranks[0]["nz"] == single["stats"]["assembled"]

    def total(launches, kind, key, pred):
        return sum(l[key] for l in launches if l["kind"] == kind and pred(l))

    sub = lambda l: l["level"] >= depth      # noqa: E731
    top = lambda l: l["level"] < depth       # noqa: E731
    for kind, key in (("gemm_grouped", "flops"), ("potrf_tile", "ctas"), ("trsm_tile", "ctas")):
        # subtree levels: split with no overlap
        assert sum(total(r["launches"], kind, key, sub) for r in ranks) == pytest.approx(
            total(single["launches"], kind, key, sub), rel=1e-12)
    # top levels: small kernels replicated on every rank, large GEMM launches split by tiles
    for r in ranks:
        assert total(r["launches"], "potrf_tile", "ctas", top) == total(single["launches"], "potrf_tile", "ctas", top)
        assert total(r["launches"], "trsm_tile", "ctas", top) == total(single["launches"], "trsm_tile", "ctas", top)
    repl = [total(r["launches"], "gemm_grouped", "flops", lambda l: top(l) and not l["shared"]) for r in ranks]
    assert all(x == pytest.approx(repl[0], rel=1e-12) for x in repl)
    shared = sum(total(r["launches"], "gemm_grouped", "flops", lambda l: top(l) and l["shared"]) for r in ranks)
    assert shared + repl[0] == pytest.approx(total(single["launches"], "gemm_grouped", "flops", top), rel=1e-9)
    # both split kinds are present: broadcast-stored tiles (1) and owner-local tiles (2)
    kinds = {l["shared"] for r in ranks for l in r["launches"] if l["kind"] == "gemm_grouped" and top(l)}
    assert 1 in kinds
    # one all-reduce of the top copies per rank, and a barrier after every shared launch
    for r in ranks:
        # one reduction per shared top panel, bracketed by two barriers; one barrier per broadcast launch
        assert sum(l["kind"] == "allreduce_top" for l in r["launches"]) == world - 1
        assert sum(l["kind"] == "peer_barrier" for l in r["launches"]) == r["stats"]["shared_launches"] + 2
        assert r["stats"]["top_doubles"] == ranks[0]["stats"]["top_doubles"] > 0


    # the solve schedule (solve.cc): subtree levels split with no overlap, top levels replicated, and the top
    # part of the right-hand side that crosses the ranks is exactly the rows of the shared separators
    one = single["solve"]
    assert one["top"] == {k: 0 for k in one["top"]}
    split = {k: sum(r["solve"]["subtree"][k] for r in ranks) for k in one["subtree"]}
    tops = [r["solve"]["top"] for r in ranks]
    assert all(t == tops[0] for t in tops)
    # pulls of the level just below the top hit the same destination cluster from several subtrees, so the
    # split schedules may hold more pull slabs than the single one; everything else adds up exactly
    for k in one["subtree"]:
        if k == "pull":
            assert split[k] + tops[0][k] >= one["subtree"][k]
        else:
            assert split[k] + tops[0][k] == one["subtree"][k], k
    nsep = len(single["sizes"])
    top_rows = sum(single["sizes"][nsep - (world - 1):])   # the world - 1 highest labels are the shared separators
    assert all(r["top_size"] == top_rows for r in ranks) and top_rows > 0


def test_partition_rejects_bad_world():
    from cholesky_b200 import Cholesky, CholeskyError
    with pytest.raises(CholeskyError):
        Cholesky().set_partition(0, 3)
    with pytest.raises(CholeskyError):
        Cholesky().generate(3, 3, 1, 5, 2).set_partition(0, 4).analyze()
