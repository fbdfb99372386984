mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/pytest_gpu.log; cat gpurun_out/pytest_gpu.log
timeout 300 python tools/launch_report.py --workload lapl3d_7pt_128 > gpurun_out/launch_report_128.md 2> gpurun_out/lr.err; head -3 gpurun_out/launch_report_128.md | cut -c1-300
