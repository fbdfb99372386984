mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_8.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -q -k "(distinct or partitioned) and (8- or 4-)" 2>&1 | tail -15 > gpurun_out/pytest_mgpu_8.log
cat gpurun_out/pytest_mgpu_8.log
for g in 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $g --steps 3 --warmup 2 > gpurun_out/bench128_${g}gpu.json 2> gpurun_out/bench128_${g}gpu.err
  cat gpurun_out/bench128_${g}gpu.json; tail -2 gpurun_out/bench128_${g}gpu.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 3 --warmup 2 --workload lapl3d_27pt_96 > gpurun_out/bench_27pt_96_8gpu.json 2> gpurun_out/bench_27pt_96_8gpu.err
cat gpurun_out/bench_27pt_96_8gpu.json; tail -2 gpurun_out/bench_27pt_96_8gpu.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/launch_report.py --workload lapl3d_7pt_128 > gpurun_out/launch_report_128_8gpu.md 2> gpurun_out/lr_8.err
head -12 gpurun_out/launch_report_128_8gpu.md
