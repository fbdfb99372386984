#!/usr/bin/env python
"""One factorization and two solves of a named workload (for ncu captures of the solve kernels)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import WORKLOADS  # noqa: E402
from cholesky_b200 import Cholesky  # noqa: E402

ch = Cholesky(0).generate(*WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "lapl3d_7pt_64"]).analyze()
ch.factor()
b = np.ones(ch.n)
ch.solve(b)
x = ch.solve(b)
print("residual", float(np.linalg.norm(b - ch.matvec(x)) / np.linalg.norm(b)))
