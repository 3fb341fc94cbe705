#!/bin/bash
# Round-2 GPU call: usage scripts/gpu_r2.sh <tag> [what...]; what in: tests smoke c2 c4 c5 ref launches ncu_mlp micro sanit
TAG=${1:-r2}; shift
WHAT=${@:-tests smoke c2 c4}
O=gpurun_out
mkdir -p $O
has() { [[ " $WHAT " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_gpu.txt 2>&1
if has tests; then
  timeout 1500 python -m pytest tests -m gpu -q -s -x > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/${TAG}_pytest.log
  grep -E "passed|failed|error" $O/${TAG}_pytest.log | tail -5
  grep -E "^\[|vs oracle|rel-L2" $O/${TAG}_pytest.log | head -80
fi
if has testsall; then
  timeout 1500 python -m pytest tests -m gpu -q -s > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/${TAG}_pytest.log
  grep -E "passed|failed|error|FAILED|Error" $O/${TAG}_pytest.log | tail -40
  grep -E "^\[|vs oracle" $O/${TAG}_pytest.log | head -80
fi
if has smoke; then
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke exit $?" | tee -a $O/${TAG}_smoke.log; tail -5 $O/${TAG}_smoke.log
fi
if has c2; then
  timeout 600 python bench.py > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err; echo "bench c2 exit $?"; cat $O/${TAG}_bench_c2.json; tail -3 $O/${TAG}_bench_c2.err
fi
if has c4; then
  timeout 600 python bench.py --config c4 --steps 5 --warmup 3 > $O/${TAG}_bench_c4_x3.json 2> $O/${TAG}_bench_c4_x3.err; echo "bench c4 x3 exit $?"; cat $O/${TAG}_bench_c4_x3.json; tail -3 $O/${TAG}_bench_c4_x3.err
  timeout 600 python bench.py --config c4 --steps 5 --warmup 3 --precision bf16 --no-cpu-baseline > $O/${TAG}_bench_c4_bf16.json 2> $O/${TAG}_bench_c4_bf16.err; echo "bench c4 bf16 exit $?"; cat $O/${TAG}_bench_c4_bf16.json; tail -3 $O/${TAG}_bench_c4_bf16.err
fi
if has c5; then
  timeout 600 python bench.py --config c5 --steps 10 --no-micro > $O/${TAG}_bench_c5.json 2> $O/${TAG}_bench_c5.err; echo "bench c5 exit $?"; cat $O/${TAG}_bench_c5.json; tail -3 $O/${TAG}_bench_c5.err
fi
if has sweep; then
  for side in 12 16 24 32 48; do
    NIW_OVERLAP_SIDE_CTAS=$side timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-micro > $O/${TAG}_sweep_$side.json 2> $O/${TAG}_sweep_$side.err
    python -c "import json,sys; d=json.load(open('$O/${TAG}_sweep_$side.json')); print('side_ctas $side: ms/step %.4f value %.0f e2e %.0f mlp_ms %.4f' % (d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['mlp_ms_per_step']))"
  done
fi
if has ref; then
  timeout 900 python bench.py --impl reference --steps 20 --warmup 5 > $O/${TAG}_bench_ref.json 2>&1; cat $O/${TAG}_bench_ref.json
fi
if has micro; then
  timeout 300 python scripts/micro_hbm.py --json $O/${TAG}_micro_hbm.json 2>&1 | tee $O/${TAG}_micro_hbm.log
fi
if has launches; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-micro > $O/${TAG}_launches.log 2>&1
  python scripts/summarize_launches.py $O/${TAG}_launches.csv > $O/${TAG}_launches.md 2>&1; tail -40 $O/${TAG}_launches.md
fi
if has ncu_mlp; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_(fwd|dx|dw)_kernel' -s 6 -c 3 -f \
    -o $O/${TAG}_mlp python scripts/ncu_mlp.py 1024 > $O/${TAG}_ncu_mlp.log 2>&1
fi
if has sanit; then
  timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_tc.py -q -x -k "not full_size and not c2" > $O/${TAG}_memcheck_tc.log 2>&1; echo "memcheck exit $?" | tee -a $O/${TAG}_memcheck_tc.log
  tail -15 $O/${TAG}_memcheck_tc.log
  timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_tc.py -q -x -k "selftest or x3_forward or (forward_vs_fp32 and 37)" > $O/${TAG}_racecheck_tc.log 2>&1; echo "racecheck exit $?" | tee -a $O/${TAG}_racecheck_tc.log
  tail -15 $O/${TAG}_racecheck_tc.log
fi
ls -la $O | tail -30
