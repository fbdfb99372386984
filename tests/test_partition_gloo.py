"""CPU tests of the multi-GPU partition (world_size 2 and 4 over the gloo backend): every rank
compiles its own schedule; together they must cover the single-rank schedule exactly once -- the
subtree levels by subtree, the top levels by owned row blocks.  No GPU, no compute calls.
(tests/test_schedule_sim.py executes such schedules on the host and checks the factor.)"""
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
    os.environ["CHOL_ROW_BLOCK"] = "64"   # deal the rows of this small grid's top panels in blocks of 64
    os.environ["CHOL_NBO"] = "64"         # ... and block every other panel (and the single-rank schedule) the same way
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
    # subtree levels: the Schur updates are split with no overlap (the in-panel work depends on how many rows a launch
    # holds -- few: panel_kernel slabs, many: trsm_tile + GEMM -- so it is compared through the tile counts below)
    schur = lambda l: l["level"] >= depth and l["phase"] == 4      # noqa: E731
    assert sum(total(r["launches"], "gemm_grouped", "flops", schur) for r in ranks) == pytest.approx(
        total(single["launches"], "gemm_grouped", "flops", schur), rel=1e-12)
    # top levels: every pivot tile is factored by exactly one rank (the owner of its diagonal block) ...
    assert sum(r["stats"]["diag_tiles"] for r in ranks) == single["stats"]["diag_tiles"]
    # ... and the executed flops of the top levels are those of the single-rank schedule, dealt out with no overlap
    # (a few masked rows at odd ownership boundaries and the per-tile accounting of split launches aside)
    topschur = lambda l: top(l) and l["phase"] == 4      # noqa: E731
    split_flops = sum(total(r["launches"], "gemm_grouped", "flops", topschur) for r in ranks)
    assert split_flops == pytest.approx(total(single["launches"], "gemm_grouped", "flops", topschur), rel=3e-2)
    per_rank = [total(r["launches"], "gemm_grouped", "flops", top) for r in ranks]
    assert max(per_rank) <= 1.35 * min(per_rank)          # dealt evenly (small grid: coarse blocks)
    for r in ranks:
        kinds = [l["kind"] for l in r["launches"]]
        # partial sums of the rows a rank owns: one reduction per top level; two world barriers around them and one
        # at the end of every top level
        assert kinds.count("reduce_rects") == depth
        assert r["stats"]["push_launches"] > 0
        assert r["stats"]["top_doubles"] == ranks[0]["stats"]["top_doubles"] > 0
        # streams: the chain of the diagonal blocks on 1, the rows of the top panels on 3, background pushes on 2,
        # trailing and Schur updates and the reductions on 0
        assert {l["stream"] for l in r["launches"] if l["kind"] == "panel_kernel"} == {1, 3}
        assert {l["stream"] for l in r["launches"] if l["kind"] == "panel_kernel" and l["level"] >= depth} == {1}
        assert {l["stream"] for l in r["launches"] if l["kind"] == "reduce_rects"} == {0}
    nsync = [sum(l["kind"] == "peer_sync" for l in r["launches"]) for r in ranks]
    assert min(nsync) >= 2 + depth

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
