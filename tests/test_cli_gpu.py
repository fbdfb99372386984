"""The reference's own integration tests (test_matrices.py:49-142), with `regent.py mmat.rg ...` swapped for
the drop-in command line `cholesky_b200/cholesky` (same flags, mmat.rg:1072-1093).  Each case: run the
program on a fixture, then apply verify.check_matrix / verify.check_solution's comparisons (1e-4) against
the golden vectors that verify.py itself produced (tests/golden/make_golden.py)."""
import os
import subprocess

import numpy as np
import pytest

from conftest import CASES

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CLI = os.path.join(ROOT, "cholesky_b200", "cholesky")


def run(**kw):
    """test_matrices.py:23-35"""
    args = [CLI, "-i", kw["mat"], "-s", kw["separators"], "-c", kw["clusters"], "-b", kw["b"], "-o", kw["solution"],
            "-m", kw["factored_mat"]]
    with open(kw["stdout"], "w") as f:
        return subprocess.call(args, stdout=f)


@pytest.mark.parametrize("case", CASES)
def test_matrices(case, golden, tmp_path):
    import scipy.io
    g = golden[case]
    out = tmp_path / "output"
    out.mkdir()
    args = dict(mat=g.mtx, separators=g.ord, clusters=g.clust, b=g.b, factored_mat=str(out / "factored.mtx"),
                solution=str(out / "solution.mtx"), stdout=str(out / "stdout"))
    assert run(**args) == 0, open(args["stdout"]).read()
    # verify.check_matrix (verify.py:278-287): factor file vs scipy.linalg.cholesky(permute_matrix(...))
    regent = np.tril(np.asarray(scipy.io.mmread(args["factored_mat"]).todense()))
    assert np.allclose(g.L_dense(), regent, rtol=1e-04, atol=1e-04)
    # verify.check_solution (verify.py:290-302): solution file vs scipy.linalg.solve(A, b)
    sol = np.genfromtxt(args["solution"]).reshape(-1)
    assert np.allclose(g.x, sol, rtol=1e-04, atol=1e-04)
    # the progress lines the reference prints (mmat.rg:1095-1121)
    text = open(args["stdout"]).read()
    s = g.struct
    for line in ("Iterations: 1", f"M: {s['n']} N: {s['n']} nz: {s['nz']}", f"levels: {s['levels']}",
                 f"separators: {s['nsep']}", f"Blocks ispace: {s['blocks']}", f"Clusters ispace: {s['clusters0']}"):
        assert line in text, line


def test_devices_flag_runs_the_partitioned_path(golden, tmp_path):
    """--devices 0,0: two ranks driven from the one process (here sharing cuda:0), same outputs as the single-GPU run"""
    import scipy.io
    g = golden["lapl_3375x3375"]
    out = tmp_path / "output"
    out.mkdir()
    args = [CLI, "-i", g.mtx, "-s", g.ord, "-c", g.clust, "-b", g.b, "-o", str(out / "solution.mtx"), "-m", str(out / "factored.mtx"),
            "--devices", "0,0"]
    env = dict(os.environ, CHOL_ROW_BLOCK="64", CUDA_DEVICE_MAX_CONNECTIONS="32")
    with open(out / "stdout", "w") as f:
        assert subprocess.call(args, stdout=f, env=env) == 0, open(out / "stdout").read()
    assert "GPUs: 2" in open(out / "stdout").read()
    L = np.tril(np.asarray(scipy.io.mmread(str(out / "factored.mtx")).todense()))
    assert np.allclose(g.L_dense(), L, rtol=1e-04, atol=1e-04)
    assert np.allclose(g.x, np.genfromtxt(str(out / "solution.mtx")).reshape(-1), rtol=1e-04, atol=1e-04)
    assert subprocess.call([CLI, "-i", g.mtx, "-s", g.ord, "-c", g.clust, "--gpus", "3"], stdout=subprocess.DEVNULL) != 0


def test_iterations_flag(golden, tmp_path):
    g = golden["lapl_400x400"]
    out = str(tmp_path / "stdout")
    with open(out, "w") as f:
        rc = subprocess.call([CLI, "-i", g.mtx, "-s", g.ord, "-c", g.clust, "--iterations", "3"], stdout=f)
    assert rc == 0
    assert "Iterations: 3" in open(out).read()
