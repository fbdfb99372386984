"""CPU tests: the oracle (oracle/chol_oracle.c) against the golden vectors produced by the
reference's own verify.py (tests/golden/make_golden.py).  These pin the oracle; the GPU parity
tests then compare the CUDA path with the pinned oracle."""
import os

import numpy as np
import pytest

from conftest import CASES, entrywise_ok
from oracle import oracle as orc


@pytest.fixture(scope="module")
def oracles(golden):
    out = {}
    for c in CASES:
        g = golden[c]
        o = orc.Oracle(g.mtx, g.ord, g.clust, literal_assembly=1)
        o.factor(threads=1)
        out[c] = o
    return out


@pytest.mark.parametrize("case", CASES)
def test_structure_matches_golden(case, golden, oracles):
    g, o = golden[case], oracles[case]
    s = g.struct
    assert (o.n, o.nz, o.levels, o.num_separators) == (s["n"], s["nz"], s["levels"], s["nsep"])
    assert o.num_blocks() == s["blocks"]
    assert o.num_clusters0() == s["clusters0"]
    assert [o.num_filled(t) for t in range(o.levels)] == s["filled"]
    assert o.call_counts() == s["calls"]
    assert o.factor_nnz() == s["nnzL"]


@pytest.mark.parametrize("case", CASES)
def test_permutation_matches_verify_py(case, golden, oracles):
    """permuted A assembled by the oracle == verify.permute_matrix (independent tree/permutation code)"""
    g = golden[case]
    o = orc.Oracle(g.mtx, g.ord, g.clust, literal_assembly=1)
    o.assemble()
    np.testing.assert_array_equal(o.factor_dense(), g.pmat_dense())


@pytest.mark.parametrize("case", CASES)
def test_factor_matches_scipy_golden(case, golden, oracles):
    g, o = golden[case], oracles[case]
    L = o.factor_dense()
    Lref = g.L_dense()
    ok, worst = entrywise_ok(L, Lref, rtol=1e-12)
    assert ok, worst
    # same structural nonzeros as the dense factor
    I, J, V = o.factor_coo()
    assert set(zip(I.tolist(), J.tolist())) == set(zip(g.L[0].tolist(), g.L[1].tolist()))
    A = g.pmat_dense()
    A = A + np.tril(A, -1).T
    assert np.linalg.norm(A - L @ L.T) / np.linalg.norm(A) <= 1e-12


@pytest.mark.parametrize("case", CASES)
def test_literal_and_scatter_assembly_agree(case, golden, oracles):
    g = golden[case]
    a = oracles[case]
    b = orc.Oracle(g.mtx, g.ord, g.clust, literal_assembly=0)
    b.factor(threads=1)
    for t in range(a.levels):
        np.testing.assert_array_equal(a.filled(t), b.filled(t))
        assert a.filled_checksum(t) == b.filled_checksum(t)
    np.testing.assert_array_equal(a.factor_dense(), b.factor_dense())


@pytest.mark.parametrize("case", CASES)
def test_threaded_factor_agrees(case, golden, oracles):
    g = golden[case]
    b = orc.Oracle(g.mtx, g.ord, g.clust)
    b.factor(threads=4)
    ok, worst = entrywise_ok(b.factor_dense(), oracles[case].factor_dense(), rtol=1e-12)
    assert ok, worst


@pytest.mark.parametrize("case", CASES)
def test_solve_matches_golden(case, golden, oracles):
    g, o = golden[case], oracles[case]
    b = orc.read_vector(g.b, g.n)
    x = o.solve(b)
    assert np.allclose(x, g.x, rtol=1e-10, atol=1e-10)


@pytest.mark.parametrize("case", CASES)
def test_factor_file_roundtrip_passes_reference_tolerance(case, golden, oracles, tmp_path):
    """write_matrix format (%0.8g) read back with scipy.io.mmread, as verify.check_matrix does"""
    import scipy.io
    g, o = golden[case], oracles[case]
    p = str(tmp_path / "factor.mtx")
    o.write_factor(p)
    with open(p) as f:
        assert f.readline().strip() == "%%MatrixMarket matrix coordinate real hermitian"
        assert f.readline().split() == [str(g.n), str(g.n), str(g.struct["nnzL"])]
    # mmread symmetrises a hermitian file; verify.check_matrix then takes np.tril
    L = np.tril(np.asarray(scipy.io.mmread(p).todense()))
    assert np.allclose(g.L_dense(), L, rtol=1e-4, atol=1e-4)


def test_hash_sax_known_values():
    """uthash HASH_SAX over the 8 little-endian key bytes (uthash.h:602-610), restated in Python"""
    def sax(key):
        h = 0
        for b in key.to_bytes(8, "little"):
            h ^= ((h << 5) + (h >> 2) + b) & 0xFFFFFFFFFFFFFFFF
        return h
    for k in (0, 1, 24, 3375 * 3374 + 17, 2**40 + 12345):
        assert orc.hash_sax(k) == sax(k)


def test_mmio_ref_banner_if_built(golden):
    """oracle/_ref/libmmio_ref.so is the reference's own mmio.c compiled where it lies; when present,
    its banner/size parse must agree with the oracle's restatement"""
    import ctypes as C
    so = os.path.join(os.path.dirname(orc.HERE), "oracle", "_ref", "libmmio_ref.so")
    if not os.path.exists(so):
        pytest.skip("oracle/_ref not built (reference tree absent)")
    ref = C.CDLL(so)
    libc = C.CDLL(None)
    libc.fopen.restype = C.c_void_p
    libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
    libc.fclose.argtypes = [C.c_void_p]
    ref.mm_read_banner.argtypes = [C.c_void_p, C.c_char_p]
    ref.mm_read_mtx_crd_size.argtypes = [C.c_void_p] + [C.POINTER(C.c_int)] * 3
    for c in CASES:
        g = golden[c]
        f = libc.fopen(g.mtx.encode(), b"r")
        tc = C.create_string_buffer(4)
        assert ref.mm_read_banner(f, tc) == 0
        M, N, nz = C.c_int(), C.c_int(), C.c_int()
        assert ref.mm_read_mtx_crd_size(f, C.byref(M), C.byref(N), C.byref(nz)) == 0
        libc.fclose(f)
        assert tc.raw == b"MCRH"
        assert (M.value, N.value, nz.value) == (g.n, g.n, g.struct["nz"])


@pytest.mark.parametrize("grid", [(12, 12, 12, 27, 0), (11, 9, 7, 7, 4), (40, 33, 1, 5, 0)])
def test_oracle_on_generated_grids_matches_dense_cholesky(grid, tmp_path):
    """beyond the fixtures: the generated inputs of the BASELINE stencils (27-point is this repo's definition,
    SURVEY 8(d)) through the oracle against SciPy's dense factor of the permuted matrix, and the solve"""
    import scipy.io
    import scipy.linalg
    from cholesky_b200 import Cholesky
    m, o, c = (str(tmp_path / x) for x in ("a.mtx", "a_ord.txt", "a_clust.txt"))
    ch = Cholesky().generate(*grid)
    ch.write_inputs(m, o, c)
    orc_ = orc.Oracle(m, o, c)
    orc_.factor(threads=1)
    A = np.asarray(scipy.io.mmread(m).todense())
    perm = orc_.perm()
    Lref = scipy.linalg.cholesky(A[np.ix_(perm, perm)], lower=True)
    L = orc_.factor_dense()
    scale = np.maximum(np.abs(Lref), 1e-6 * np.abs(Lref).max())
    assert np.max(np.abs(L - Lref) / scale) <= 1e-11
    assert np.count_nonzero(L) == orc_.factor_nnz()
    b = np.random.default_rng(0).integers(1, 11, size=A.shape[0]).astype(np.float64)
    x = orc_.solve(b)
    assert np.linalg.norm(A @ x - b) / np.linalg.norm(b) <= 1e-12
