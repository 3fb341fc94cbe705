#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tc.py tests/test_gpu_graph.py tests/test_gpu_engine.py -x -q -m gpu > gpurun_out/dwred_tests.log 2>&1
tail -4 gpurun_out/dwred_tests.log
run() { timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-micro 2>/dev/null | tail -1 | python -c "import json,sys;d=json.loads(sys.stdin.read());print('$1', d['ms_per_step'], d['value'], d['roofline']['frac'])"; }
NIW_OVERLAP_SIDE_CTAS=17 NIW_DW_PARTIALS=1 run "partials side=17"
for s in 20 17 13 11; do NIW_OVERLAP_SIDE_CTAS=$s run "direct side=$s"; done
NIW_OVERLAP_SIDE_CTAS=17 NIW_DW_PARTIALS=1 run "partials side=17"
NIW_OVERLAP_SIDE_CTAS=17 timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-micro --timeline gpurun_out/dwred_tl > /dev/null 2>&1
