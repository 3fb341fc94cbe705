for ts in 1 0; do
  echo "== NIW_P2P_TWO_SHOT=$ts"
  NIW_P2P_TWO_SHOT=$ts timeout 400 python -m pytest tests/test_gpu_multi.py -q -x -s -k "p2p" 2>&1 | grep -E "passed|failed|rel-L2|Error" | head
  NIW_P2P_TWO_SHOT=$ts timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-loss-check > gpurun_out/r2y_n2_ts$ts.json 2> gpurun_out/r2y_n2_ts$ts.err
  python -c "import json; d=json.load(open('gpurun_out/r2y_n2_ts$ts.json')); print('N=2 two_shot=$ts ms/step %.4f value %.0f' % (d['ms_per_step'], d['value']), d.get('collective_errors'))" || tail -5 gpurun_out/r2y_n2_ts$ts.err
done
