"""GPU parity tests: the CUDA numeric factorization (through the C ABI) against the pinned CPU
oracle and the golden vectors of the reference's verify.py.  Tolerances (BASELINE.json north_star):
factor entries within 1e-10 relative with an absolute floor (conftest.entrywise_ok), relative
residual <= 1e-12, pattern bit-exact."""
import numpy as np
import pytest

from conftest import CASES, compare_coo, entrywise_ok
from cholesky_b200 import Cholesky, read_vector
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def _coo_dict(I, J, V):
    return {(int(i), int(j)): float(v) for i, j, v in zip(I, J, V)}


@pytest.mark.parametrize("case", CASES)
def test_factor_matches_oracle_and_golden(case, golden):
    g = golden[case]
    ch = Cholesky().load(g.mtx, g.ord, g.clust).analyze(keep_records=True)
    st = ch.factor()
    assert st.info == 0 and st.kernel_launches > 0
    o = orc.Oracle(g.mtx, g.ord, g.clust)
    o.factor(threads=1)
    L, Lo, Lg = ch.factor_dense(), o.factor_dense(), g.L_dense()
    ok, worst = entrywise_ok(L, Lo)
    assert ok, f"vs oracle {worst}"
    ok, worst = entrywise_ok(L, Lg)
    assert ok, f"vs scipy golden {worst}"
    # identical structural nonzeros
    assert ch.factor_nnz() == g.struct["nnzL"]
    I, J, _ = ch.factor_coo()
    Io, Jo, _ = o.factor_coo()
    assert set(zip(I.tolist(), J.tolist())) == set(zip(Io.tolist(), Jo.tolist()))
    A = g.pmat_dense()
    A = A + np.tril(A, -1).T
    assert np.linalg.norm(A - L @ L.T) / np.linalg.norm(A) <= 1e-12
    assert ch.residual(k=8) <= 1e-12


@pytest.mark.parametrize("case", CASES)
def test_reference_acceptance_checks(case, golden, tmp_path):
    """what test_matrices.py asserts: check_matrix on the written factor and check_solution, both 1e-4"""
    import scipy.io
    g = golden[case]
    ch = Cholesky().load(g.mtx, g.ord, g.clust).analyze()
    ch.factor()
    p = str(tmp_path / "factor.mtx")
    ch.write_factor(p)
    L = np.tril(np.asarray(scipy.io.mmread(p).todense()))
    assert np.allclose(g.L_dense(), L, rtol=1e-4, atol=1e-4)
    x = ch.solve(read_vector(g.b, g.n))
    assert np.allclose(x, g.x, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("case", ["lapl_400x400", "lapl_3375x3375"])
def test_piecewise_fused_tasks_match_oracle(case, golden):
    """level by level, phase by phase: fused_dpotrf / fused_dtrsm / fused_dsyrk+dgemm"""
    g = golden[case]
    ch = Cholesky().load(g.mtx, g.ord, g.clust).analyze()
    o = orc.Oracle(g.mtx, g.ord, g.clust)
    ch.assemble()
    o.assemble()
    np.testing.assert_array_equal(ch.factor_dense(), o.factor_dense())  # assembly is exact
    for lvl in range(ch.levels - 1, -1, -1):
        for step in ("fused_dpotrf", "fused_dtrsm", "fused_update"):
            getattr(ch, step)(lvl)
            getattr(o, step)(lvl)
            ok, worst = entrywise_ok(ch.factor_dense(), o.factor_dense())
            assert ok, f"level {lvl} {step}: {worst}"


@pytest.mark.parametrize("grid", [(16, 16, 16, 7, 0), (33, 17, 1, 5, 0), (12, 12, 12, 27, 0), (24, 20, 9, 7, 6),
                                  (40, 40, 1, 5, 3), (30, 30, 30, 7, 4)])
def test_generated_grids_match_oracle(grid, tmp_path):
    ch = Cholesky().generate(*grid)
    m, o_, c = (str(tmp_path / x) for x in ("a.mtx", "a_ord.txt", "a_clust.txt"))
    ch.write_inputs(m, o_, c)
    ch.analyze()
    ch.factor()
    o = orc.Oracle(m, o_, c)
    o.factor(threads=2)
    I, J, V = ch.factor_coo()
    Io, Jo, Vo = o.factor_coo()
    a, b = _coo_dict(I, J, V), _coo_dict(Io, Jo, Vo)
    assert a.keys() == b.keys()
    ref = np.array([b[k] for k in a])
    got = np.array([a[k] for k in a])
    scale = np.maximum(np.abs(ref), 1e-6 * np.abs(ref).max())
    assert np.max(np.abs(got - ref) / scale) <= 1e-10
    assert ch.residual(k=8) <= 1e-12
    # solve on the GPU (forward / backward sweeps over the panels) against the oracle's dtrsv/dgemv sweep
    b = np.random.default_rng(0).integers(1, 11, size=ch.n).astype(np.float64)
    x, xo = ch.solve(b), o.solve(b)
    assert np.max(np.abs(x - xo)) <= 1e-10 * np.max(np.abs(xo))


def test_large_front_blocking_paths(tmp_path):
    """a 2-level tree over a 40x40x40 grid has a 1600-dof root: exercises the left-looking outer
    blocks (NBO) and the odd-sized last tile"""
    ch = Cholesky().generate(41, 40, 39, 7, 2).analyze()
    ch.factor()
    assert ch.residual(k=4) <= 1e-12


@pytest.mark.parametrize("fused_rows_max", ["0", "1000000"])
def test_both_paths_for_the_rows_below_a_diagonal_block(monkeypatch, fused_rows_max):
    """the rows below the diagonal blocks of a launch go through panel_kernel's slabs when they are few and through
    trsm_tile + grouped GEMM launches when they are many (schedule.cc); force each path for every launch"""
    monkeypatch.setenv("CHOL_FUSED_ROWS_MAX", fused_rows_max)
    ch = Cholesky().generate(41, 40, 39, 7, 3).analyze()
    ch.factor()
    assert ch.residual(k=4) <= 1e-12
    b = np.random.default_rng(1).integers(1, 11, size=ch.n).astype(np.float64)
    x = ch.solve(b)
    assert np.linalg.norm(b - ch.matvec(x)) <= 1e-12 * np.linalg.norm(b)


def test_idempotent_and_deterministic(golden):
    g = golden["lapl_3375x3375"]
    ch = Cholesky().load(g.mtx, g.ord, g.clust).analyze()
    ch.factor()
    a = ch.factor_coo()
    ch.factor(iterations=2)
    b = ch.factor_coo()
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)  # bit-identical run to run (atomic-free accumulation)


def test_end_to_end_host_call(golden):
    g = golden["lapl_3375x3375"]
    ch = Cholesky().load(g.mtx, g.ord, g.clust).analyze()
    diag, st = ch.factor_host()
    Ld = np.diag(g.L_dense())
    assert np.allclose(diag, Ld, rtol=1e-10, atol=0)


def test_not_positive_definite_is_reported(golden, tmp_path):
    from cholesky_b200 import CholeskyError
    g = golden["lapl_25x25"]
    bad = tmp_path / "neg.mtx"
    bad.write_text(open(g.mtx).read().replace("1 1 4.0", "1 1 -4.0", 1))
    ch = Cholesky().load(str(bad), g.ord, g.clust).analyze()
    with pytest.raises(CholeskyError, match="not positive definite"):
        ch.factor()


def _entrywise_vs_oracle(grid, tmp_path, threads=8, solve_tol=1e-12):
    """factor a generated grid on the GPU and with the CPU oracle; compare every stored entry and the solve"""
    m, o, c = (str(tmp_path / x) for x in ("a.mtx", "a_ord.txt", "a_clust.txt"))
    ch = Cholesky().generate(*grid)
    ch.write_inputs(m, o, c)
    ch.analyze()
    st = ch.factor()
    assert st.info == 0
    assert ch.residual(k=4) <= 1e-12
    ref = orc.Oracle(m, o, c)
    ref.factor(threads=threads)
    same, worst = compare_coo(ch.n, ch.factor_coo(), ref.factor_coo())
    assert same, "structural nonzeros differ from the oracle's"
    assert worst <= 1e-10, worst
    b = np.random.default_rng(0).integers(1, 11, size=ch.n).astype(np.float64)
    x = ch.solve(b)
    assert np.linalg.norm(b - ch.matvec(x)) <= solve_tol * np.linalg.norm(b)
    return ch


def test_config2_512x512_entrywise(tmp_path):
    """BASELINE config 2 at full size (262 144 unknowns, 1.9e7 factor entries): every entry against the oracle"""
    # (the solve residual scales with the condition number: ~n for the 2-D Laplacian, 6e-12 measured with either factor)
    _entrywise_vs_oracle((512, 512, 1, 5, 0), tmp_path, solve_tol=1e-10)


def test_config3_64cubed_entrywise(tmp_path):
    """BASELINE config 3 at full size (262 144 unknowns, 1.6e8 factor entries): every entry against the oracle"""
    _entrywise_vs_oracle((64, 64, 64, 7, 0), tmp_path)


def test_config5_stencil_entrywise(tmp_path):
    """the 27-point stencil of BASELINE config 5 on a 40^3 grid (larger fronts per unknown): every entry against the oracle"""
    _entrywise_vs_oracle((40, 40, 40, 27, 0), tmp_path)


def test_solve_and_results_are_refused_after_a_failed_factorization(golden, tmp_path):
    from cholesky_b200 import CholeskyError
    g = golden["lapl_25x25"]
    bad = tmp_path / "neg.mtx"
    bad.write_text(open(g.mtx).read().replace("1 1 4.0", "1 1 -4.0", 1))
    ch = Cholesky().load(str(bad), g.ord, g.clust).analyze()
    with pytest.raises(CholeskyError, match="not positive definite"):
        ch.factor()
    with pytest.raises(CholeskyError, match="factor first"):
        ch.solve(np.ones(ch.n))
    with pytest.raises(CholeskyError, match="factor first"):
        ch.residual()


def test_reload_invalidates_device_state(golden):
    """a handle that is reloaded must not serve results of (or index into) the previous problem"""
    from cholesky_b200 import CholeskyError
    g, g2 = golden["lapl_3375x3375"], golden["lapl_400x400"]
    ch = Cholesky().load(g.mtx, g.ord, g.clust).analyze()
    ch.factor()
    ch.load(g2.mtx, g2.ord, g2.clust)
    with pytest.raises(CholeskyError):
        ch.solve(np.ones(g2.n))
    with pytest.raises(CholeskyError):
        ch.factor_nnz()
    ch.analyze()
    ch.factor()
    ok, worst = entrywise_ok(ch.factor_dense(), g2.L_dense())
    assert ok, worst


@pytest.mark.parametrize("case", ["lapl_400x400", "lapl_3375x3375"])
def test_binary_factor_dump_round_trip(case, golden, tmp_path):
    """row f-4: the binary block dump and its streaming converter give the file write_matrix would"""
    import scipy.io
    from cholesky_b200 import factor_binary_to_mtx, read_factor_binary
    g = golden[case]
    ch = Cholesky().load(g.mtx, g.ord, g.clust).analyze()
    ch.factor()
    b, t, u = str(tmp_path / "f.bin"), str(tmp_path / "f.mtx"), str(tmp_path / "direct.mtx")
    ch.write_factor_binary(b)
    n, I, J, V = read_factor_binary(b)
    Ic, Jc, Vc = ch.factor_coo()
    assert n == g.n and _coo_dict(I, J, V) == _coo_dict(Ic, Jc, Vc)
    factor_binary_to_mtx(b, t)
    ch.write_factor(u)
    assert sorted(open(t).read().splitlines()[2:]) == sorted(open(u).read().splitlines()[2:])
    assert open(t).read().splitlines()[:2] == open(u).read().splitlines()[:2]
    L = np.tril(np.asarray(scipy.io.mmread(t).todense()))
    assert np.allclose(g.L_dense(), L, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("case", ["lapl_25x25", "lapl_3375x3375"])
def test_solve_pair_equals_solve_on_a_single_gpu_handle(case, golden):
    """chol_solve_forward / chol_solve_backward (the partitioned-handle entry points): with one rank there
    is no shared top, the pair is chol_solve bit for bit and agrees with the oracle's dtrsv/dgemv sweep"""
    g = golden[case]
    ch = Cholesky().load(g.mtx, g.ord, g.clust).analyze()
    ch.factor()
    b = read_vector(g.b, g.n)
    assert ch.solve_top_size() == 0
    top = ch.solve_forward(b)
    assert top.size == 0
    x2 = ch.solve_backward(top)
    assert np.array_equal(x2, ch.solve(b))
    o = orc.Oracle(g.mtx, g.ord, g.clust)
    o.factor(threads=1)
    assert np.max(np.abs(x2 - o.solve(b))) <= 1e-10 * np.max(np.abs(x2))
