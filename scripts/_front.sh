#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_graph.py tests/test_gpu_engine.py -x -q -m gpu > gpurun_out/front_tests.log 2>&1
tail -4 gpurun_out/front_tests.log
run() { timeout 300 python bench.py --steps 40 --warmup 5 --no-cpu-baseline --no-micro 2>/dev/null | tail -1 | python -c "import json,sys;d=json.loads(sys.stdin.read());print('$1', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['frac'])"; }
run "default"
run "default"
NIW_FUSED_RAYS=1 run "fused rays"
timeout 300 python bench.py --steps 10 --warmup 5 --no-cpu-baseline --no-micro --timeline gpurun_out/front_tl > /dev/null 2>&1
timeout 400 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "nvp" > gpurun_out/front_memcheck.log 2>&1; tail -3 gpurun_out/front_memcheck.log
timeout 500 compute-sanitizer --tool racecheck python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "nvp_backward_on_a_capped or nvp_golden" > gpurun_out/front_racecheck.log 2>&1; tail -3 gpurun_out/front_racecheck.log
