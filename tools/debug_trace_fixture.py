#!/usr/bin/env python
"""Run the drop-in command line with `-d` on one reference fixture (materialised from tests/golden/fixtures.npz)
and leave the trace -- stdout log, per-task snapshots, factor -- in a directory, ready for the reference's
own verify.debug_factor (verify.py:216-275).  Needs a GPU.
  python tools/debug_trace_fixture.py lapl_400x400 gpurun_out/debug_lapl_400x400"""
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    case, out = sys.argv[1], sys.argv[2]
    os.makedirs(out, exist_ok=True)
    z = np.load(os.path.join(ROOT, "tests", "golden", "fixtures.npz"))
    paths = {}
    for kind in ("mtx", "ord", "clust"):
        paths[kind] = os.path.join(out, str(z[f"{case}/name/{kind}"]))
        with open(paths[kind], "wb") as f:
            f.write(z[f"{case}/file/{kind}"].tobytes())
    with open(os.path.join(out, "log.txt"), "w") as f:
        rc = subprocess.call([os.path.join(ROOT, "cholesky_b200", "cholesky"), "-i", paths["mtx"], "-s", paths["ord"],
                              "-c", paths["clust"], "-m", os.path.join(out, "factored.mtx"), "-d", out], stdout=f)
    print(case, "exit", rc, "files", len(os.listdir(out)))
    sys.exit(rc)


if __name__ == "__main__":
    main()
