#!/bin/bash
# compute-sanitizer over the NVP tests (rebuilt backward kernel, bulk-copied weights): scripts/gpu_sanit_nvp.sh <tag>
TAG=${1:-r2}; O=gpurun_out; mkdir -p $O
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "nvp" > $O/${TAG}_memcheck_nvp.log 2>&1; echo "memcheck exit $?" | tee -a $O/${TAG}_memcheck_nvp.log
tail -4 $O/${TAG}_memcheck_nvp.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_ops.py -x -q -m gpu -k "nvp_backward_on_a_capped or nvp_golden or nvp_vs_oracle" > $O/${TAG}_racecheck_nvp.log 2>&1; echo "racecheck exit $?" | tee -a $O/${TAG}_racecheck_nvp.log
tail -4 $O/${TAG}_racecheck_nvp.log
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_engine.py -x -q -m gpu -k "one_launch or prepack" > $O/${TAG}_memcheck_rays.log 2>&1; echo "memcheck exit $?" | tee -a $O/${TAG}_memcheck_rays.log
tail -4 $O/${TAG}_memcheck_rays.log
