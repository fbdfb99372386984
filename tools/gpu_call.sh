mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_multi.py -x -q -k "solve or golden or acceptance or config or group" 2>&1 | tail -3
for pf in 1 0; do echo "== prefetch $pf"; CHOL_SOLVE_PREFETCH=$pf CHOL_SOLVE_TIMES=1 python tools/solve_step.py lapl3d_7pt_128 2>&1 | grep "tile_fwd\|residual" | head -2 | cut -c1-200; done
python - <<'PY'
import sys, time
sys.path.insert(0, '.')
import numpy as np
from cholesky_b200 import Cholesky
for g in [(128,128,128,7,0),(64,64,64,7,0),(512,512,1,5,0)]:
    ch = Cholesky(0).generate(*g).analyze(); ch.factor()
    b = np.random.default_rng(0).integers(1, 11, size=ch.n).astype(np.float64)
    ch.solve(b); t=time.time(); x = ch.solve(b); print(g, "solve wall ms", (time.time()-t)*1e3, "residual", np.linalg.norm(b - ch.matvec(x))/np.linalg.norm(b))
PY
