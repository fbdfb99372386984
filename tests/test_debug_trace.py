"""The `-d` debug path: log grammar and per-task snapshots, CPU side.
 * the oracle's trace replays cleanly (tests/debug_replay.py restates verify.debug_factor);
 * the engine's host-side log is byte-identical to the oracle's (two independent implementations);
 * when the reference tree is mounted, its UNMODIFIED verify.debug_factor accepts every group of the
   oracle's trace up to the root transition it cannot get past by construction (see debug_replay.py)."""
import contextlib
import io
import os
import sys

import numpy as np
import pytest

from cholesky_b200.engine import Cholesky
from oracle.oracle import Oracle
from debug_replay import parse_log, replay

SMALL = ["lapl_9x9", "lapl_25x25", "lapl_400x400"]
# sha256 of the `-d` log of each fixture, recorded when the oracle's trace of the same fixtures was accepted by
# the reference's unmodified verify.debug_factor (profiles/debug_trace_r01.md): a change of the log grammar
# or of the task order has to be deliberate
LOG_SHA256 = {
    "lapl_9x9": "73b4a3a16ee5fd7c8e640111425f57cd4401565a52f1d52504a9ff43b7a6ed83",
    "lapl_25x25": "655b326601348c5bb965e9a0833bdf3bbca00b7ba6f42823a7977369e824e09a",
    "lapl_400x400": "a8735baf437b6731e30629315e12539a82e5b957566c78e2466c83690fca5c8f",
    "lapl_3375x3375": "18751c9b1db4f61e1642854bf317a9869637181f4b2586a9d0b842e799bcec69",
}
REF = "/root/reference"


@pytest.fixture(scope="module")
def traces(golden, tmp_path_factory):
    out = {}
    for case in SMALL:
        fx = golden[case]
        d = str(tmp_path_factory.mktemp("dbg_" + case))
        o = Oracle(fx.mtx, fx.ord, fx.clust)
        o.debug_trace(d, os.path.join(d, "log.txt"), full_precision=True)
        o.write_factor(os.path.join(d, "factored.mtx"))
        out[case] = (d, o)
    return out


@pytest.mark.parametrize("case", SMALL)
def test_oracle_trace_replays(case, golden, traces):
    fx, (d, o) = golden[case], traces[case]
    checked, files, worst, mat = replay(fx.pmat_dense(), os.path.join(d, "log.txt"), d, rtol=1e-9, atol=1e-11)
    calls = fx.struct["calls"]
    _, _, ops = parse_log(os.path.join(d, "log.txt"))
    assert sum(l["op"] == "POTRF" for l in ops) == calls["potrf"]
    assert sum(l["op"] == "TRSM" for l in ops) == calls["trsm"]
    assert sum(l["op"] == "GEMM" for l in ops) == calls["syrk"] + calls["gemm"]
    assert checked > 0 and len(set(files)) == len(files)
    # the replayed matrix ends as the factor (verify.py:272)
    assert np.allclose(np.tril(mat), fx.L_dense(), rtol=1e-9, atol=1e-11)
    # one snapshot per fused task of the level loop: sum over separators of 1 + depth + depth (depth + 1) / 2
    L = fx.struct["levels"]
    want = sum((1 << l) * (1 + l + l * (l + 1) // 2) for l in range(L))
    assert len([f for f in os.listdir(d) if "_lvl" in f and f.endswith(".mtx")]) == want


@pytest.mark.parametrize("case", SMALL + ["lapl_3375x3375"])
def test_engine_log_is_the_oracle_log(case, golden, traces, tmp_path):
    fx = golden[case]
    ch = Cholesky().load(fx.mtx, fx.ord, fx.clust).analyze(keep_records=True)
    mine = str(tmp_path / "engine.log")
    ch.write_debug_log(mine)
    import hashlib
    assert hashlib.sha256(open(mine, "rb").read()).hexdigest() == LOG_SHA256[case]
    blocks, clusters, ops = parse_log(mine)
    assert len(blocks) == fx.struct["blocks"]
    calls = fx.struct["calls"]
    assert [sum(l["op"] == k for l in ops) for k in ("POTRF", "TRSM", "GEMM")] == \
        [calls["potrf"], calls["trsm"], calls["syrk"] + calls["gemm"]]
    for l in ops:  # sizes agree with the rectangles, Level and Interval label are tied (mmat.rg:1228, 1349)
        for b in "ABC":
            if b in l:
                size = l.get(f"Size{b}", l.get(f"size{b}"))
                assert size == (l[f"{b}_Hi"][0] - l[f"{b}_Lo"][0] + 1, l[f"{b}_Hi"][1] - l[f"{b}_Lo"][1] + 1)
        assert l["Interval"] == fx.struct["levels"] - 1 - l["Level"]
    if case in traces:
        with open(os.path.join(traces[case][0], "log.txt")) as f, open(mine) as g:
            assert f.read() == g.read()


def test_debug_log_needs_records(golden):
    fx = golden["lapl_9x9"]
    ch = Cholesky().load(fx.mtx, fx.ord, fx.clust).analyze(keep_records=False)
    with pytest.raises(RuntimeError, match="keep_records"):
        ch.write_debug_log(os.devnull)


@pytest.mark.skipif(not os.path.exists(os.path.join(REF, "verify.py")), reason="reference tree not mounted")
@pytest.mark.parametrize("case", ["lapl_9x9", "lapl_25x25"])
def test_unmodified_verify_debug_factor_accepts_the_oracle_trace(case, golden, traces):
    fx, (d, o) = golden[case], traces[case]
    sys.path.insert(0, REF)
    try:
        import verify
    finally:
        sys.path.remove(REF)
    buf = io.StringIO()
    raised = False
    with contextlib.redirect_stdout(buf):
        try:
            verify.debug_factor(fx.mtx, fx.ord, os.path.join(d, "factored.mtx"), os.path.join(d, "log.txt"), d)
        except AssertionError:
            raised = True
    seen = [l.split()[-1] for l in buf.getvalue().splitlines() if l.startswith("Verifying:")]
    _, files, _, _ = replay(fx.pmat_dense(), os.path.join(d, "log.txt"), d)
    # every group before the root's POTRF is compared and accepted; the comparison that raises is the
    # last Schur update into the root block, made after the root POTRF was already applied
    assert raised and seen == files[:len(seen)] and len(seen) == len(files) - 1
    N = fx.struct["nsep"]
    assert seen[-1].endswith(f"_c{N}{N}.mtx") and "lvl1" in seen[-1]


@pytest.mark.parametrize("grid", [(3, 3, 1, 5, 3), (5, 4, 3, 7, 4), (7, 1, 1, 5, 2), (9, 9, 9, 27, 4), (1, 1, 1, 5, 1)])
def test_generated_grids_engine_log_equals_oracle_log_and_replays(grid, tmp_path):
    """beyond the four fixtures: ragged grids, one-dof separators, a single-separator tree, the 27-point stencil"""
    import scipy.io
    m, o, c = (str(tmp_path / x) for x in ("a.mtx", "a_ord.txt", "a_clust.txt"))
    ch = Cholesky().generate(*grid)
    ch.write_inputs(m, o, c)
    ch.analyze(keep_records=True)
    ch.write_debug_log(str(tmp_path / "engine.log"))
    d = str(tmp_path / "trace")
    orc = Oracle(m, o, c)
    orc.debug_trace(d, str(tmp_path / "oracle.log"), full_precision=True)
    assert open(str(tmp_path / "engine.log")).read() == open(str(tmp_path / "oracle.log")).read()
    A = np.asarray(scipy.io.mmread(m).todense())
    perm = ch.perm()
    checked, files, worst, mat = replay(np.tril(A[np.ix_(perm, perm)]), str(tmp_path / "oracle.log"), d, rtol=1e-9, atol=1e-11)
    assert checked == len(files) >= 1
    assert np.allclose(np.tril(mat), orc.factor_dense(), rtol=1e-9, atol=1e-11)
