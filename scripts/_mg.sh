N=${1:-2}
timeout 300 python bench.py --steps 30 --no-cpu-baseline --no-micro > gpurun_out/r2t_bench_n1.json 2> gpurun_out/r2t_bench_n1.err
python -c "import json; d=json.load(open('gpurun_out/r2t_bench_n1.json')); print('N=1 ms/step %.4f value %.0f e2e %.0f' % (d['ms_per_step'], d['value'], d['e2e']['value']))" || tail -5 gpurun_out/r2t_bench_n1.err
for p2p in 1 0; do
  NIW_P2P_ALLREDUCE=$p2p timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --timeline gpurun_out/r2t_tl_n${N}_p2p$p2p > gpurun_out/r2t_bench_n${N}_p2p$p2p.json 2> gpurun_out/r2t_bench_n${N}_p2p$p2p.err
  python -c "import json; d=json.load(open('gpurun_out/r2t_bench_n${N}_p2p$p2p.json')); print('N=$N p2p=$p2p ms/step %.4f value %.0f e2e %.0f' % (d['ms_per_step'], d['value'], d['e2e']['value']))" || tail -5 gpurun_out/r2t_bench_n${N}_p2p$p2p.err
done
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_engine.py -q -x 2>&1 | tail -3
