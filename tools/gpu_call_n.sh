# usage: bash tools/gpu_call_n.sh "<gpu counts>" [workload]
mkdir -p gpurun_out
W=${2:-lapl3d_7pt_128}
for g in $1; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $g --steps 3 --warmup 2 --workload $W 2> gpurun_out/bench_${W}_${g}gpu.err | grep '^{' > gpurun_out/bench_${W}_${g}gpu.json
  python -c "
import json; d=json.load(open('gpurun_out/bench_${W}_${g}gpu.json')); print('$W', $g, round(d['ms_per_step'],2), round(d['value']), d['factor']['residual'], d['factor']['solve_rel_residual'], d['factor']['top_copies_max_diff'], round(d['factor']['solve_ms'],1), round(d['factor']['analyze_s'],1), d['roofline']['kernel_ms'], d.get('INVALID'))"
done
