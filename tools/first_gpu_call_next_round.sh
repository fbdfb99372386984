#!/bin/bash
# One GPU call that decides the queued, default-off variants: parity + timing of each, the GPU suite with the
# two measured winners switched on, and the headline workload under every variant.
mkdir -p gpurun_out
python tools/check_experimental.py > gpurun_out/experimental.jsonl 2> gpurun_out/experimental.err
CHOL_POTRF_R=2 CHOL_TRSM_BATCH=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/pytest_potrf_r2_trsm_batch.log
for v in "X=0" "CHOL_POTRF_R=2 CHOL_TRSM_BATCH=1" "CHOL_POTRF_R=3 CHOL_TRSM_BATCH=1" "CHOL_GEMM_STAGES=4" "CHOL_GRAPH=1"; do
  echo "== $v"
  env $v python tools/profile_step.py --workload lapl3d_7pt_128 --iterations 3 --warmup 1 | tail -1
done > gpurun_out/variants_128.log 2>&1
cat gpurun_out/experimental.jsonl gpurun_out/pytest_potrf_r2_trsm_batch.log gpurun_out/variants_128.log
