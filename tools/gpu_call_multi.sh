# usage: bash tools/gpu_call_multi.sh N   (N GPUs visible)
N=$1
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus_$N.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -q -k "distinct or partitioned" 2>&1 | tail -15 > gpurun_out/pytest_mgpu_$N.log
cat gpurun_out/pytest_mgpu_$N.log
for n in $(seq 1 4); do
  g=$((1 << n)); [ $g -le $N ] || break
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $g --steps 3 --warmup 2 > gpurun_out/bench128_${g}gpu.json 2> gpurun_out/bench128_${g}gpu.err
  cat gpurun_out/bench128_${g}gpu.json; tail -2 gpurun_out/bench128_${g}gpu.err
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 tools/launch_report.py --workload lapl3d_7pt_128 > gpurun_out/launch_report_128_${N}gpu.md 2> gpurun_out/lr_$N.err
head -40 gpurun_out/launch_report_128_${N}gpu.md
