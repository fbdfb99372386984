#!/usr/bin/env python
"""Feed a `-d` trace produced on the GPU (tools/debug_trace_fixture.py) to the reference's UNMODIFIED
verify.debug_factor and verify.check_matrix.  Runs only where the reference tree is mounted (build container).
  python tools/check_debug_trace_with_reference.py lapl_400x400 gpurun_out/debug_lapl_400x400"""
import contextlib
import io
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, "/root/reference")
import verify  # noqa: E402  (the reference's own file)
from debug_replay import replay  # noqa: E402


def main():
    case, d = sys.argv[1], sys.argv[2]
    z = np.load(os.path.join(ROOT, "tests", "golden", "fixtures.npz"))
    mtx, o = (os.path.join(d, str(z[f"{case}/name/{k}"])) for k in ("mtx", "ord"))
    buf, raised = io.StringIO(), False
    with contextlib.redirect_stdout(buf):
        try:
            verify.debug_factor(mtx, o, os.path.join(d, "factored.mtx"), os.path.join(d, "log.txt"), d)
        except AssertionError:
            raised = True
    seen = [l.split()[-1] for l in buf.getvalue().splitlines() if l.startswith("Verifying:")]
    n = int(z[f"{case}/pmat/I"].max()) + 1
    pm = np.zeros((n, n))
    pm[z[f"{case}/pmat/I"], z[f"{case}/pmat/J"]] = z[f"{case}/pmat/V"]
    checked, files, worst, _ = replay(pm, os.path.join(d, "log.txt"), d)
    ok, _, _ = verify.check_matrix(mtx, o, os.path.join(d, "factored.mtx"))
    print(f"{case}: unmodified verify.debug_factor accepted {len(seen) - (1 if raised else 0)} of {len(files)} task groups"
          + (f", raised at {seen[-1]}" if raised else "")
          + f"; compare-then-apply replay accepted {checked} (worst abs diff {worst:.2e}); verify.check_matrix {ok}")


if __name__ == "__main__":
    main()
