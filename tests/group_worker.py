"""Worker of the group-handle tests (own process: CUDA_DEVICE_MAX_CONNECTIONS has to be set before CUDA starts):
factor a generated grid on a group of ranks driven from ONE process through the C ABI -- chol_create(devices, ngpu)
-- and compare with the CPU oracle.  `devices` may repeat a GPU (several ranks share it), so the partitioned path
is exercised on a single-GPU box too.
  python tests/group_worker.py nx,ny,nz,stencil,levels d0,d1,..."""
import json
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from conftest import compare_coo  # noqa: E402

from cholesky_b200 import Cholesky  # noqa: E402


def main():
    grid = tuple(int(x) for x in sys.argv[1].split(","))
    devices = [int(x) for x in sys.argv[2].split(",")]
    from oracle import oracle as orc
    ch = Cholesky(devices=devices).generate(*grid)
    tmp = tempfile.mkdtemp()
    m, o, c = (os.path.join(tmp, x) for x in ("a.mtx", "a_ord.txt", "a_clust.txt"))
    ch.write_inputs(m, o, c)
    ch.analyze()
    st = ch.factor(iterations=2, warmup=1)
    out = dict(world=len(devices), grid=grid, seconds=st.seconds_best, launches=int(st.kernel_launches), info=int(st.info))
    out["copies_diff"] = ch.top_copies_diff()
    out["residual"] = ch.residual(k=4)
    rhs = np.random.default_rng(0).integers(1, 11, size=ch.n).astype(np.float64)
    x = ch.solve(rhs)
    out["solve_residual"] = float(np.linalg.norm(rhs - ch.matvec(x)) / np.linalg.norm(rhs))
    diag, _ = ch.factor_host()   # the end-to-end entry with host buffers
    entries = ch.n <= 200000      # beyond that: size-independent properties only (the COO lists get too long)
    if entries:
        ref = orc.Oracle(m, o, c)
        ref.factor(threads=4)
        Io, Jo, Vo = ref.factor_coo()
        out["pattern_equal"], out["worst_entry"] = compare_coo(ch.n, ch.factor_coo(), (Io, Jo, Vo))
        xo = ref.solve(rhs)
        out["solve_vs_oracle"] = float(np.max(np.abs(x - xo)) / np.max(np.abs(xo)))
        dsel = Io == Jo
        out["diag_worst"] = float(np.max(np.abs(diag[Io[dsel]] - Vo[dsel]) / np.abs(Vo[dsel])))
    else:
        out.update(pattern_equal=True, worst_entry=0.0, solve_vs_oracle=0.0, diag_worst=0.0, entries_checked=False)
    out["push_launches"] = [ch.rank_handle(r).partition_stats()["push_launches"] for r in range(ch.num_ranks())] if len(devices) > 1 else [0]
    out["ok"] = bool(out["pattern_equal"] and out["worst_entry"] <= 1e-10 and out["copies_diff"] == 0.0 and out["residual"] <= 1e-12
                     and out["solve_vs_oracle"] <= 1e-10 and out["solve_residual"] <= 1e-12 and out["diag_worst"] <= 1e-10 and out["info"] == 0)
    print(json.dumps(out))
    sys.exit(0 if out["ok"] else 1)


if __name__ == "__main__":
    main()
