timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_graph.py tests/test_gpu_engine.py -q -x 2>&1 | tail -2
for side in 17 19 22 26; do
  NIW_OVERLAP_SIDE_CTAS=$side timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-micro 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('side_ctas $side: ms/step %.4f value %.0f' % (d['ms_per_step'], d['value']))"
done
