"""The `-d` debug path on the GPU: the level loop one fused task group at a time, one
snapshot per reference task.  Checked three ways: the replay that restates verify.debug_factor accepts
the trace; every snapshot agrees with the oracle's on the block its task wrote (1e-10 entry-wise); and the
factor the handle is left with is the regular one."""
import os
import subprocess

import numpy as np
import pytest

from conftest import entrywise_ok
from cholesky_b200 import Cholesky
from debug_replay import compare_traces, replay
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "cholesky_b200", "cholesky")


@pytest.mark.parametrize("case", ["lapl_9x9", "lapl_25x25", "lapl_400x400"])
def test_debug_trace_matches_oracle_trace(case, golden, tmp_path):
    g = golden[case]
    mine, ref = str(tmp_path / "gpu"), str(tmp_path / "oracle")
    os.makedirs(mine), os.makedirs(ref)
    ch = Cholesky().load(g.mtx, g.ord, g.clust).analyze(keep_records=True)
    log = os.path.join(mine, "log.txt")
    ch.write_debug_log(log)
    ch.factor_debug(mine, full_precision=True, with_txt=(case != "lapl_400x400"))
    o = orc.Oracle(g.mtx, g.ord, g.clust)
    o.debug_trace(ref, os.path.join(ref, "log.txt"), full_precision=True)
    assert open(log).read() == open(os.path.join(ref, "log.txt")).read()
    names = sorted(f for f in os.listdir(ref) if f.endswith(".mtx"))
    assert sorted(f for f in os.listdir(mine) if f.endswith(".mtx")) == names
    checked, files, worst, mat = replay(g.pmat_dense(), log, mine, rtol=1e-9, atol=1e-11)
    assert checked == len(files) > 0
    assert compare_traces(log, mine, ref) == checked
    # the handle holds the complete factor afterwards, and a regular factorization still works on it
    ok, w = entrywise_ok(ch.factor_dense(), g.L_dense())
    assert ok, w
    ch.factor()
    ok, w = entrywise_ok(ch.factor_dense(), g.L_dense())
    assert ok, w
    if case != "lapl_400x400":  # the "%0.2f" block dumps of write_blocks (mmat.rg:185-217)
        txt = open(os.path.join(mine, names[0][:-4] + ".txt")).read().splitlines()
        assert txt[0].startswith("Level: ") and txt[1].startswith("Color: ")


def test_cli_debug_flag(golden, tmp_path):
    """`-d <dir>` as the reference (mmat.rg:1086-1090): log on stdout, snapshots under the directory"""
    g = golden["lapl_25x25"]
    d = tmp_path / "dbg"
    d.mkdir()
    out = str(tmp_path / "stdout")
    fac = str(tmp_path / "factored.mtx")
    with open(out, "w") as f:
        rc = subprocess.call([CLI, "-i", g.mtx, "-s", g.ord, "-c", g.clust, "-m", fac, "-d", str(d)], stdout=f)
    assert rc == 0, open(out).read()
    checked, files, worst, _ = replay(g.pmat_dense(), out, str(d))  # "%0.8g" files: the reference's 1e-4
    assert checked == len(files) > 0
    import scipy.io
    L = np.tril(np.asarray(scipy.io.mmread(fac).todense()))
    assert np.allclose(g.L_dense(), L, rtol=1e-4, atol=1e-4)
