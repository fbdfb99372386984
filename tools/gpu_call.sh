mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 600 python tools/time_workloads.py > gpurun_out/time_workloads.log 2>&1; cat gpurun_out/time_workloads.log
