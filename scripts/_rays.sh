#!/bin/bash
mkdir -p gpurun_out
for f in 1 0 1 0 1 0; do
  NIW_FUSED_RAYS=$f timeout 600 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-micro 2>/dev/null | tail -1 > gpurun_out/rays_bench_$f.json
  python -c "import json;d=json.load(open('gpurun_out/rays_bench_$f.json'));print('fused=$f', d['ms_per_step'], d['value'], d['e2e']['value'], d['gpu_launches'], d['clocks'])"
done
