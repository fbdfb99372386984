"""The multi-GPU schedule, executed on the host (tests/sim/sched_sim.cc: an interpreter of the compiled launch
lists with one factor buffer per rank, adversarially random interleavings that respect only stream order, the
list's events and the flag words) and compared entry by entry with the CPU oracle.  What this pins without a
GPU: ownership of the top panels' rows, the pushes and partial-sum reductions, and that no wait is missing
(a missing one gives a wrong factor for some seed, a cyclic one a deadlock report)."""
import os
import sys
import tempfile

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from conftest import entrywise_ok  # noqa: E402
from sim import sim  # noqa: E402

from cholesky_b200 import Cholesky  # noqa: E402
from oracle import oracle as orc  # noqa: E402

_ORACLE = {}


def oracle_factor(grid):
    if grid not in _ORACLE:
        tmp = tempfile.mkdtemp()
        m, o, c = (os.path.join(tmp, x) for x in ("a.mtx", "a_ord.txt", "a_clust.txt"))
        Cholesky().generate(*grid).write_inputs(m, o, c)
        ref = orc.Oracle(m, o, c)
        ref.factor(threads=2)
        _ORACLE[grid] = ref.factor_dense()
    return _ORACLE[grid]


@pytest.mark.parametrize("world", [1, 2, 4, 8])
@pytest.mark.parametrize("grid,row_block", [((12, 12, 12, 7, 4), 64), ((33, 31, 5, 7, 5), 64), ((40, 40, 1, 5, 4), 64),
                                            ((9, 9, 9, 27, 4), 64), ((16, 16, 16, 7, 5), 128)])
def test_simulated_partitioned_factor_matches_oracle(monkeypatch, world, grid, row_block):
    monkeypatch.setenv("CHOL_ROW_BLOCK", str(row_block))
    Lr = oracle_factor(grid)
    for seed in (1, 2, 3, 4, 5):
        L, copy_diff, st = sim.factor(grid, world, seed)
        ok, worst = entrywise_ok(L, Lr)
        assert ok, (world, seed, worst)
        assert copy_diff == 0.0          # every rank ends with bit-identical copies of the top panels
        if world > 1:
            assert st["push_rects"] > 0 and st["reduces"] > 0


@pytest.mark.parametrize("world", [1, 4])
def test_simulated_schedule_with_the_throughput_path_for_the_rows(monkeypatch, world):
    """CHOL_FUSED_ROWS_MAX=0: the rows below every diagonal block go through trsm_tile + grouped GEMM launches per
    64-column tile step instead of panel_kernel's slabs (the path large fronts take by default)"""
    monkeypatch.setenv("CHOL_ROW_BLOCK", "128")
    monkeypatch.setenv("CHOL_FUSED_ROWS_MAX", "0")
    grid = (16, 16, 16, 7, 5)
    L, copy_diff, _ = sim.factor(grid, world, 7)
    ok, worst = entrywise_ok(L, oracle_factor(grid))
    assert ok and copy_diff == 0.0, worst


@pytest.mark.parametrize("world", [2, 8])
def test_simulated_top_levels_with_the_throughput_path_for_the_rows(monkeypatch, world):
    """CHOL_FUSED_ROWS_MAX_TOP=0: on the top panels of a partition the rows below the diagonal blocks go through
    trsm_tile + grouped GEMM launches on the rows stream (what a rank does when it owns more than two waves of slabs)"""
    monkeypatch.setenv("CHOL_ROW_BLOCK", "64")
    monkeypatch.setenv("CHOL_FUSED_ROWS_MAX_TOP", "0")
    grid = (20, 20, 20, 7, 5)
    for seed in (1, 2, 3):
        L, copy_diff, _ = sim.factor(grid, world, seed)
        ok, worst = entrywise_ok(L, oracle_factor(grid))
        assert ok and copy_diff == 0.0, worst


@pytest.mark.parametrize("world", [2, 8])
def test_simulated_schedule_without_lookahead(monkeypatch, world):
    """CHOL_LOOKAHEAD=0 puts every launch of a rank on one stream in list order: the order of the instrumented
    (per-launch timing) pass.  It must not deadlock and must give the same factor."""
    monkeypatch.setenv("CHOL_ROW_BLOCK", "64")
    monkeypatch.setenv("CHOL_LOOKAHEAD", "0")
    grid = (12, 12, 12, 7, 4)
    L, copy_diff, _ = sim.factor(grid, world, 5)
    ok, worst = entrywise_ok(L, oracle_factor(grid))
    assert ok and copy_diff == 0.0, worst
