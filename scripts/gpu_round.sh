#!/bin/bash
# One gpurun call: GPU parity tests, smoke, bench lines (C2 + C5-per-GPU size), reference arm, HBM micro-bench,
# ncu launch list and ncu --set full captures.  Everything lands in gpurun_out/<tag>_*.
# usage: scripts/gpu_round.sh <tag> [what...]   what in: tests bench micro launches ncu_mlp ncu_hbm  (default: all)
TAG=${1:-r1}; shift
WHAT=${@:-tests bench micro launches ncu_mlp ncu_hbm}
O=gpurun_out
mkdir -p $O
has() { [[ " $WHAT " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/${TAG}_gpu.txt 2>&1
if has tests; then
  timeout 900 python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest.log 2>&1; echo "pytest exit $?" | tee -a $O/${TAG}_pytest.log
  tail -5 $O/${TAG}_pytest.log
  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/${TAG}_smoke.log 2>&1; echo "smoke exit $?" | tee -a $O/${TAG}_smoke.log
fi
if has bench; then
  timeout 600 python bench.py > $O/${TAG}_bench_c2.json 2> $O/${TAG}_bench_c2.err; echo "bench exit $?"; cat $O/${TAG}_bench_c2.json
  timeout 600 python bench.py --rays 8192 --no-cpu-baseline > $O/${TAG}_bench_8192.json 2> $O/${TAG}_bench_8192.err; cat $O/${TAG}_bench_8192.json
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/${TAG}_bench_ref.json 2>&1; cat $O/${TAG}_bench_ref.json
fi
if has micro; then
  timeout 300 python scripts/micro_hbm.py --json $O/${TAG}_micro_hbm.json 2>&1 | tee $O/${TAG}_micro_hbm.log
fi
if has launches; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file $O/${TAG}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-micro > $O/${TAG}_launches.log 2>&1
fi
if has ncu_mlp; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:'tc_(fwd|dx|dw)_kernel' -s 6 -c 3 -f \
    -o $O/${TAG}_mlp python scripts/ncu_mlp.py 1024 > $O/${TAG}_ncu_mlp.log 2>&1
fi
if has ncu_hbm; then
  timeout 600 ncu --set full --clock-control none --import-source on \
    -k regex:'composite_|stratified_kernel|pdf_merge|raygen_pose' -c 8 -f \
    -o $O/${TAG}_hbm python scripts/micro_hbm.py --once > $O/${TAG}_ncu_hbm.log 2>&1
fi
ls -la $O | tail -30
