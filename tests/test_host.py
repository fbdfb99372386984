"""CPU tests of the product's host side: file readers, generators, symbolic analysis and schedule
compiler, compared bit for bit with the pinned oracle.  No compute calls on a GPU here."""
import ctypes as C
import filecmp
import os

import numpy as np
import pytest

from conftest import CASES
from cholesky_b200 import Cholesky, CholeskyError, _lib
from oracle import oracle as orc

GRIDS = [(15, 15, 15, 7, 5), (16, 16, 16, 7, 0), (33, 17, 1, 5, 0), (12, 12, 12, 27, 0), (24, 20, 9, 7, 6)]


def test_library_exports_every_declared_symbol():
    L = _lib.load()
    for name in _lib.EXPORTS:
        assert hasattr(L, name), name
    # and the headers declare nothing that is not exported
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    import re
    declared = set()
    for h in ("cholesky.h", "chol_mmio.h", "chol_mnd.h"):
        src = open(os.path.join(root, "include", h)).read()
        src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
        declared |= set(re.findall(r"\b((?:chol|mm|mnd)_[a-z0-9_]+|register_mappers)\s*\(", src))
    declared = {d for d in declared if not d.startswith(("mm_is_", "mm_set_", "mm_clear", "mm_initialize"))}
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)


@pytest.mark.parametrize("case", CASES)
def test_symbolic_matches_oracle_on_reference_fixtures(case, golden):
    g = golden[case]
    ch = Cholesky().load(g.mtx, g.ord, g.clust).analyze(keep_records=True)
    o = orc.Oracle(g.mtx, g.ord, g.clust)
    s = g.struct
    assert (ch.n, ch.nz, ch.levels, ch.num_separators) == (s["n"], s["nz"], s["levels"], s["nsep"])
    assert ch.num_blocks() == s["blocks"] and ch.num_clusters0() == s["clusters0"]
    assert [ch.num_filled(t) for t in range(ch.levels)] == s["filled"]
    assert ch.call_counts() == s["calls"]
    np.testing.assert_array_equal(ch.perm(), o.perm())
    np.testing.assert_array_equal(ch.sep_sizes(), o.sep_sizes())
    np.testing.assert_array_equal(ch.block_bounds(), o.block_bounds())
    assert ch.max_int_size() == o.max_int_size()
    for t in range(ch.levels):
        np.testing.assert_array_equal(ch.filled(t), o.filled(t))
        assert ch.filled_checksum(t) == o.filled_checksum(t)
    assert ch.flops() == o.flops()
    fa, fb = ch.flops_by_level(), o.flops_by_level()
    for k in fa:
        np.testing.assert_array_equal(fa[k], fb[k])


@pytest.mark.parametrize("grid,name", [((3, 3, 1, 5, 2), "lapl_9x9"), ((5, 5, 1, 5, 3), "lapl_25x25"),
                                       ((20, 20, 1, 5, 5), "lapl_400x400"), ((15, 15, 15, 7, 5), "lapl_3375x3375")])
def test_generated_laplacian_is_byte_identical_to_fixture(grid, name, golden, tmp_path):
    ch = Cholesky().generate(*grid)
    m = str(tmp_path / "a.mtx")
    ch.write_inputs(m, None, None)
    assert filecmp.cmp(m, golden[name].mtx, shallow=False)


@pytest.mark.parametrize("grid", GRIDS)
def test_generated_nd_files_roundtrip_and_match_oracle(grid, tmp_path):
    ch = Cholesky().generate(*grid)
    m, o, c = (str(tmp_path / x) for x in ("a.mtx", "a_ord.txt", "a_clust.txt"))
    ch.write_inputs(m, o, c)
    ch.analyze(keep_records=True)
    ch2 = Cholesky().load(m, o, c).analyze(keep_records=True)   # through the text readers
    oc = orc.Oracle(m, o, c)
    assert ch.levels == oc.levels and ch.num_separators == oc.num_separators
    np.testing.assert_array_equal(ch.perm(), oc.perm())
    np.testing.assert_array_equal(ch2.perm(), oc.perm())
    for t in range(ch.levels):
        np.testing.assert_array_equal(ch.filled(t), oc.filled(t))
        np.testing.assert_array_equal(ch2.filled(t), oc.filled(t))
        assert ch.filled_checksum(t) == oc.filled_checksum(t)
    assert ch.call_counts() == oc.call_counts()
    assert ch.flops() == oc.flops()
    # separators are true vertex separators: nothing of A is dropped
    oc.assemble()
    assert oc.factor_nnz() == ch.nz


def test_random_grids_symbolic_parity(tmp_path):
    """ragged shapes: 40 seeded random grids / stencils / tree depths, pattern bit for bit against the oracle"""
    rng = np.random.default_rng(7)
    done = 0
    while done < 40:
        three_d = rng.random() < 0.6
        nx, ny = int(rng.integers(3, 14)), int(rng.integers(3, 14))
        nz = int(rng.integers(2, 10)) if three_d else 1
        st = (7 if rng.random() < 0.6 else 27) if three_d else 5
        lv = int(rng.integers(1, 5))
        try:
            ch = Cholesky().generate(nx, ny, nz, st, lv)
        except CholeskyError:          # too many levels for this grid: an empty separator is refused
            continue
        m, o, c = (str(tmp_path / x) for x in ("a.mtx", "a_ord.txt", "a_clust.txt"))
        ch.write_inputs(m, o, c)
        ch.analyze(keep_records=True)
        oc = orc.Oracle(m, o, c)
        np.testing.assert_array_equal(ch.perm(), oc.perm())
        for t in range(ch.levels):
            np.testing.assert_array_equal(ch.filled(t), oc.filled(t), err_msg=str((nx, ny, nz, st, lv, t)))
        assert ch.call_counts() == oc.call_counts() and ch.flops() == oc.flops()
        oc.assemble()
        assert oc.factor_nnz() == ch.nz, (nx, ny, nz, st, lv)      # true vertex separators: nothing dropped
        done += 1


def test_levels_rule_matches_utils_py():
    import math
    for dims, st in [((512, 512, 1), 5), ((64, 64, 64), 7)]:
        ch = Cholesky().generate(*dims, st, 0).analyze()
        n = dims[0] * dims[1] * dims[2]
        assert ch.levels == math.ceil(math.log(n / 64) / math.log(2)) + 1  # utils.py:6-7


def test_bad_inputs_are_reported(golden, tmp_path):
    g = golden["lapl_25x25"]
    with pytest.raises(CholeskyError):
        Cholesky().load(str(tmp_path / "missing.mtx"), g.ord, g.clust)
    bad = tmp_path / "bad_ord.txt"
    bad.write_text(open(g.ord).read().replace("6;6,10,12,14,18,", "6;6,10,12,14,"))
    with pytest.raises(CholeskyError):
        Cholesky().load(g.mtx, str(bad), g.clust)
    # a separator with two clusters at elimination breaks the reference's fused tasks
    badc = tmp_path / "bad_clust.txt"
    badc.write_text(open(g.clust).read().replace("0;0,4,;", "0;0,2,4,;"))
    with pytest.raises(CholeskyError):
        Cholesky().load(g.mtx, g.ord, str(badc)).analyze()


def _with_extra_entry(g, tmp_path, extra):
    """the fixture matrix with one more coordinate line (banner nz bumped)"""
    lines = open(g.mtx).read().splitlines()
    n, m, nz = lines[1].split()
    lines[1] = f"{n} {m} {int(nz) + 1}"
    p = tmp_path / "extra.mtx"
    p.write_text("\n".join(lines + [extra]) + "\n")
    return str(p)


def test_entry_between_unrelated_separators_is_dropped_like_the_reference(golden, tmp_path):
    """no block exists for two separators that are not in an ancestor relation: the reference drops the
    entry silently (the nz assert is commented out, mmat.rg:1191); pattern and flop count do not change"""
    g = golden["lapl_25x25"]
    base = Cholesky().load(g.mtx, g.ord, g.clust).analyze(keep_records=True)
    perm = base.perm()
    # permuted rows 0 and 4 belong to leaf separators 1 and 2 (4 dofs each): siblings, unrelated
    i, j = sorted((int(perm[0]) + 1, int(perm[4]) + 1), reverse=True)
    m = _with_extra_entry(g, tmp_path, f"{i} {j} -0.5")
    ch = Cholesky().load(m, g.ord, g.clust).analyze(keep_records=True)
    o = orc.Oracle(m, g.ord, g.clust)
    assert ch.nz == base.nz + 1
    for t in range(ch.levels):
        np.testing.assert_array_equal(ch.filled(t), base.filled(t))
        np.testing.assert_array_equal(ch.filled(t), o.filled(t))
    assert ch.partition_stats()["assembled"] == base.nz     # the extra entry has no destination
    assert ch.flops() == base.flops() == o.flops()


def test_explicit_zero_entry_does_not_fill(golden, tmp_path):
    """a stored zero is 'empty' to the reference's hash table (mnd.c:178-195) and to fill_block (mmat.rg:591)"""
    g = golden["lapl_25x25"]
    base = Cholesky().load(g.mtx, g.ord, g.clust).analyze(keep_records=True)
    perm = base.perm()
    i, j = sorted((int(perm[24]) + 1, int(perm[0]) + 1), reverse=True)   # (root dof, leaf dof): a real block
    m = _with_extra_entry(g, tmp_path, f"{i} {j} 0.0")
    ch = Cholesky().load(m, g.ord, g.clust).analyze(keep_records=True)
    for t in range(ch.levels):
        np.testing.assert_array_equal(ch.filled(t), base.filled(t))


def test_matrix_reader_skips_exactly_two_lines(golden, tmp_path):
    """mnd.c:162-164 skips two lines whatever they hold; a comment line therefore shifts the entries
    and the last one is lost -- the readers here and in the oracle behave the same way"""
    g = golden["lapl_9x9"]
    lines = open(g.mtx).read().splitlines()
    p = tmp_path / "commented.mtx"
    p.write_text("\n".join([lines[0], "% a comment"] + lines[1:]) + "\n")
    ch = Cholesky().load(str(p), g.ord, g.clust).analyze(keep_records=True)
    o = orc.Oracle(str(p), g.ord, g.clust)
    assert ch.nz == o.nz == 21          # the banner reader skips comments (mmio.c:189-217) ...
    for t in range(ch.levels):          # ... but the entry reader does not, and both sides agree on the result
        np.testing.assert_array_equal(ch.filled(t), o.filled(t))


def test_hash_sax_matches_oracle():
    L = _lib.load()
    for k in (0, 1, 24, 3375 * 3374 + 17, 2**40 + 12345):
        assert int(L.mnd_hash_sax(C.c_uint64(k))) == orc.hash_sax(k)


def test_numeric_path_fails_loudly_without_gpu(golden):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    g = golden["lapl_9x9"]
    ch = Cholesky().load(g.mtx, g.ord, g.clust).analyze()
    with pytest.raises(CholeskyError, match="no CUDA device"):
        ch.factor()


def test_factor_binary_converter_streams_a_dump_to_the_reference_text_format(tmp_path):
    """csrc/factor_file.cc: header + dense records -> write_matrix's "%d %d %0.8g" lines (mmat.rg:129-144)"""
    import struct
    import scipy.io
    from cholesky_b200 import CholeskyError, factor_binary_to_mtx, read_factor_binary
    rng = np.random.default_rng(3)
    n = 11
    recs = [(0, 0, 4, 4), (6, 0, 3, 4), (4, 4, 7, 7)]
    dense = np.zeros((n, n))
    body = b""
    for r0, c0, nr, nc in recs:
        v = rng.standard_normal((nr, nc))
        if r0 == c0:
            v = np.tril(v)
        v[rng.random((nr, nc)) < 0.2] = 0.0
        dense[r0:r0 + nr, c0:c0 + nc] = v
        body += struct.pack("<4i", r0, c0, nr, nc) + np.asfortranarray(v).tobytes(order="F")
    nnz = int(np.count_nonzero(dense))
    head = b"CHOLFAC1" + struct.pack("<2i", n, n) + b"MCRG" + struct.pack("<i", 0) + struct.pack("<2q", len(recs), nnz)
    p = tmp_path / "f.bin"
    p.write_bytes(head + body)
    factor_binary_to_mtx(str(p), str(tmp_path / "f.mtx"), full_precision=True)
    got = scipy.io.mmread(str(tmp_path / "f.mtx")).toarray()
    assert np.array_equal(got, dense)
    nn, I, J, V = read_factor_binary(str(p))
    back = np.zeros((n, n))
    back[I, J] = V
    assert nn == n and np.array_equal(back, dense)
    # "%0.8g" by default, as the reference
    factor_binary_to_mtx(str(p), str(tmp_path / "g.mtx"))
    assert np.allclose(scipy.io.mmread(str(tmp_path / "g.mtx")).toarray(), dense, rtol=1e-7, atol=0)
    # truncated and foreign files are refused
    (tmp_path / "t.bin").write_bytes((head + body)[:-8])
    with pytest.raises(CholeskyError, match="-4"):
        factor_binary_to_mtx(str(tmp_path / "t.bin"), str(tmp_path / "t.mtx"))
    (tmp_path / "x.bin").write_bytes(b"not a dump" * 10)
    with pytest.raises(CholeskyError, match="-2"):
        factor_binary_to_mtx(str(tmp_path / "x.bin"), str(tmp_path / "x.mtx"))


def test_level_bytes_add_up_to_the_factor_storage(golden):
    """chol_level_bytes: the panel term counts every stored entry of the level's separators once (x 16 B)"""
    fx = golden["lapl_3375x3375"]
    ch = Cholesky().load(fx.mtx, fx.ord, fx.clust).analyze()
    tot = sum(ch.level_bytes(l)["panel"] for l in range(ch.levels))
    sizes = ch.sep_sizes()
    # stored entries = nnz(L) pattern of the filled clusters: lower triangles of the pivot blocks + off-diagonal rows
    assert tot / 16 >= fx.struct["nnzL"]
    assert tot / 16 <= ch.factor_doubles()
    root = ch.level_bytes(0)
    n0 = int(sizes[-1])
    assert root["panel"] == 16.0 * n0 * (n0 + 1) / 2 and root["operands"] == 0 and root["destinations"] == 0
    leaves = ch.level_bytes(ch.levels - 1)
    assert leaves["operands"] > 0 and leaves["destinations"] > 0
    with pytest.raises(RuntimeError):
        ch.level_bytes(ch.levels)


def test_blocking_knobs_do_not_change_the_executed_work(monkeypatch):
    """the block-column width is a host decision: the launch list changes, the executed GEMM flops of the Schur
    updates and the factorization's flop total do not; narrower block columns mean more (smaller) panel launches"""
    def summary():
        ch = Cholesky().generate(24, 20, 18, 7, 5).analyze()
        ls = ch.launches()
        schur = round(sum(l["flops"] for l in ls if l["kind"] == "gemm_grouped" and l["phase"] == 4))
        return (sum(l["kind"] == "panel_kernel" for l in ls), schur, ch.flops(), ch.partition_stats()["diag_tiles"])
    base = summary()
    monkeypatch.setenv("CHOL_NBO", "128")
    a = summary()
    assert a[1:] == base[1:]
    assert a[0] > base[0]


@pytest.mark.parametrize("grid", [(512, 512, 1, 5, 0), (64, 64, 64, 7, 0), (128, 128, 128, 7, 0), (96, 96, 96, 27, 0)],
                         ids=["config2_512sq", "config3_64cubed", "config4_128cubed", "config5_96cubed_27pt"])
def test_block_pattern_is_bit_exact_at_the_baseline_sizes(grid, tmp_path):
    """BASELINE.json configs 2-5 at full size: the engine's symbolic phase against the oracle's restatement of
    compute_filled_clusters -- every `Filled` record of every interval label (count + order-independent
    checksum of all nine fields), the permutation, the reference's BLAS call counts and its flop total"""
    import shutil
    from oracle import oracle as orc
    d = str(tmp_path / "inputs")
    os.makedirs(d)
    m, o, c = (os.path.join(d, x) for x in ("a.mtx", "a_ord.txt", "a_clust.txt"))
    ch = Cholesky().generate(*grid)
    ch.write_inputs(m, o, c)
    ch.analyze()
    oc = orc.Oracle(m, o, c)
    try:
        assert ch.levels == oc.levels and ch.num_separators == oc.num_separators
        np.testing.assert_array_equal(ch.perm(), oc.perm())
        for t in range(ch.levels):
            assert ch.num_filled(t) == oc.num_filled(t), t
            assert ch.filled_checksum(t) == oc.filled_checksum(t), t
        assert ch.call_counts() == oc.call_counts()
        assert ch.flops() == oc.flops()
        assert ch.num_blocks() == oc.num_blocks() and ch.num_clusters0() == oc.num_clusters0()
    finally:
        oc.close()
        shutil.rmtree(d)


def test_analysis_file_round_trip(tmp_path):
    """chol_save_analysis / chol_load_analysis: the ranks of a node analyse once.  A handle that loads the file compiles
    the same schedule as one that analysed for itself; the file of a different problem is refused"""
    from cholesky_b200 import CholeskyError
    grid = (20, 18, 16, 7, 5)
    a = Cholesky().generate(*grid).analyze(keep_records=True)
    path = str(tmp_path / "analysis.bin")
    a.save_analysis(path)
    for rank, world in ((0, 1), (1, 2), (3, 4)):
        own = Cholesky().generate(*grid).set_partition(rank, world).analyze(keep_records=True)
        b = Cholesky().generate(*grid).set_partition(rank, world).load_analysis(path)
        assert b.launches() == own.launches()
        assert b.partition_stats() == own.partition_stats()
        assert b.flops() == own.flops() and b.call_counts() == own.call_counts() and b.factor_doubles() == own.factor_doubles()
        for t in range(b.levels):
            assert b.filled_checksum(t) == own.filled_checksum(t)
            assert np.array_equal(b.filled(t), own.filled(t))
    with pytest.raises(CholeskyError, match="different problem"):
        Cholesky().generate(20, 18, 17, 7, 5).load_analysis(path)
    (tmp_path / "junk.bin").write_bytes(b"not an analysis")
    with pytest.raises(CholeskyError):
        Cholesky().generate(*grid).load_analysis(str(tmp_path / "junk.bin"))


def test_create_refuses_unsupported_gpu_counts():
    """chol_create(devices, ngpu): 1, 2, 4 or 8 GPUs -- never silently fewer than asked for"""
    from cholesky_b200 import CholeskyError
    for bad in ([0, 1, 2], [0] * 5, [0] * 16):
        with pytest.raises(CholeskyError):
            Cholesky(devices=bad)
    g = Cholesky(devices=[0, 0, 0, 0])       # creating a group touches no device
    assert g.num_ranks() == 4
    g.generate(12, 12, 12, 7, 4).analyze()   # one symbolic analysis, four schedules
    stats = [g.rank_handle(r).partition_stats() for r in range(4)]
    assert sum(s["assembled"] for s in stats) == g.nz
