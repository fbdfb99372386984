"""Partitioned factorization on the GPU, checked entry by entry against the CPU oracle.

1. Group handles -- chol_create(devices, ngpu) drives 2 / 4 / 8 ranks from one process.  The device list may
   repeat a GPU, so these run on a single-GPU box (all ranks on cuda:0) and on distinct GPUs when the box has them.
2. One process per GPU (torchrun, CUDA-IPC peers): needs as many GPUs as ranks, skipped otherwise.
Both check: pattern identical, entries <= 1e-10 (SURVEY 7.3-9 rule), all ranks' copies of the top panels
bit-identical, GPU-side residual <= 1e-12, partitioned solve vs the oracle's dtrsv/dgemv sweep <= 1e-10."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


def _run(cmd, env, timeout=900):
    r = subprocess.run(cmd, env=env, capture_output=True, text=True, timeout=timeout)
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert lines, r.stdout[-2000:] + r.stderr[-3000:]
    out = json.loads(lines[-1])
    assert r.returncode == 0 and out["ok"], (out, r.stderr[-2000:])
    return out


GROUP_CASES = [
    # world, grid, row block (64: small panels still get several owned blocks; 256: the production value)
    (2, "30,30,30,7,4", 64), (2, "40,40,1,5,3", 64), (4, "33,31,29,7,5", 64), (8, "33,31,29,7,5", 64),
    (8, "24,24,24,27,5", 128), (4, "48,48,48,7,0", 256), (8, "64,64,64,7,0", 256),
]


@pytest.mark.parametrize("world,grid,row_block", GROUP_CASES)
def test_group_handle_matches_oracle(world, grid, row_block):
    """all ranks on cuda:0 (runs on every GPU box)"""
    if _ngpu() < 1:
        pytest.skip("needs a GPU")
    env = dict(os.environ, CHOL_ROW_BLOCK=str(row_block), CUDA_DEVICE_MAX_CONNECTIONS="32")
    out = _run([sys.executable, os.path.join(ROOT, "tests", "group_worker.py"), grid, ",".join(["0"] * world)], env)
    assert max(out["push_launches"]) > 0      # rows of the top panels were pushed to peers (a rank that owns no block of a tiny panel pushes nothing)


@pytest.mark.parametrize("world", [4, 8])
def test_group_handle_with_the_throughput_path_on_the_top_panels(world):
    """CHOL_FUSED_ROWS_MAX_TOP=0: the rows below the diagonal blocks of the top panels go through trsm_tile + grouped GEMM
    launches on the rows stream (what a rank does when it owns more than two waves of 64-row slabs of a block column)"""
    if _ngpu() < 1:
        pytest.skip("needs a GPU")
    env = dict(os.environ, CHOL_ROW_BLOCK="64", CHOL_FUSED_ROWS_MAX_TOP="0", CUDA_DEVICE_MAX_CONNECTIONS="32")
    _run([sys.executable, os.path.join(ROOT, "tests", "group_worker.py"), "33,31,29,7,5", ",".join(["0"] * world)], env)


@pytest.mark.parametrize("world,grid,row_block", [(2, "30,30,30,7,4", 64), (4, "48,48,48,7,0", 256), (8, "64,64,64,7,0", 256)])
def test_group_handle_on_distinct_gpus(world, grid, row_block):
    """one process, one rank per GPU (cudaDeviceEnablePeerAccess, no IPC)"""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, CHOL_ROW_BLOCK=str(row_block))
    _run([sys.executable, os.path.join(ROOT, "tests", "group_worker.py"), grid, ",".join(str(d) for d in range(world))], env)


@pytest.mark.parametrize("world,grid,row_block", [(2, "30,30,30,7,4", 64), (2, "40,40,1,5,3", 64), (4, "33,31,29,7,5", 64),
                                                  (8, "33,31,29,7,5", 64), (2, "48,48,48,7,0", 256), (4, "48,48,48,7,0", 256),
                                                  (8, "48,48,48,7,0", 256), (8, "64,64,64,7,0", 256)])
def test_partitioned_factor_matches_oracle(world, grid, row_block):
    """one process per GPU (torchrun), peers mapped through CUDA IPC"""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, CHOL_ROW_BLOCK=str(row_block))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "mgpu_worker.py"), grid]
    out = _run(cmd, env)
    assert max(out["push_launches"]) > 0
