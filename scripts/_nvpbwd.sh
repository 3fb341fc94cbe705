#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_graph.py tests/test_gpu_engine.py -x -q -m gpu > gpurun_out/nvpbwd_tests.log 2>&1
tail -4 gpurun_out/nvpbwd_tests.log
run() { timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-micro 2>/dev/null | tail -1 | python -c "import json,sys;d=json.loads(sys.stdin.read());print('$1', d['ms_per_step'], d['value'])"; }
for s in 22 17 13 11 10 9 8; do NIW_OVERLAP_SIDE_CTAS=$s run "side=$s"; done
NIW_OVERLAP_SIDE_CTAS=13 timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-micro --timeline gpurun_out/nvpbwd_tl > /dev/null 2>&1
