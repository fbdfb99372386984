mkdir -p gpurun_out
for g in 8 4 2; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $g --steps 3 --warmup 2 2> gpurun_out/bench128_${g}gpu.err | grep '^{' > gpurun_out/bench128_${g}gpu.json
  python -c "
import json; d=json.load(open('gpurun_out/bench128_${g}gpu.json')); print($g, d['ms_per_step'], d['value'], d['factor']['residual'], d['factor']['solve_rel_residual'], d['factor']['top_copies_max_diff'], d['factor']['analyze_s'], d['roofline']['kernel_ms'], d.get('INVALID'))"
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 8 --steps 3 --warmup 2 --workload lapl3d_27pt_96 2> gpurun_out/bench_27pt_96_8gpu.err | grep '^{' > gpurun_out/bench_27pt_96_8gpu.json
python -c "
import json; d=json.load(open('gpurun_out/bench_27pt_96_8gpu.json')); print('96^3 27pt x8', d['ms_per_step'], d['value'], d['factor']['residual'], d['factor']['solve_rel_residual'], d['factor']['top_copies_max_diff'])"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/launch_report.py --workload lapl3d_7pt_128 > gpurun_out/launch_report_128_8gpu.md 2> gpurun_out/lr_8.err
sed -n 1,18p gpurun_out/launch_report_128_8gpu.md
