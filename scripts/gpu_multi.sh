#!/bin/bash
# Multi-GPU bench lines: scripts/gpu_multi.sh <N> <config>... (run under gpurun --gpus N); c2 also writes per-rank timelines
N=${1:-2}; shift
for cfg in "$@"; do
  extra=""; [ "$cfg" = "c2" ] && extra="--timeline gpurun_out/r3x_tl_n${N}"
  steps=20; [ "$cfg" = "c5" ] && steps=10; [ "$cfg" = "c4" ] && steps=5
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --config $cfg --steps $steps --warmup 5 $extra > gpurun_out/r3x_bench_${cfg}_n${N}.json 2> gpurun_out/r3x_bench_${cfg}_n${N}.err
  python -c "import json; d=json.load(open('gpurun_out/r3x_bench_${cfg}_n${N}.json')); print('$cfg N=$N ms/step %.4f value %.0f e2e %.0f' % (d['ms_per_step'], d['value'], d['e2e']['value']), d.get('collective_errors'), (d.get('loss_check') or {}).get('abs_diff'))" || tail -5 gpurun_out/r3x_bench_${cfg}_n${N}.err
done
