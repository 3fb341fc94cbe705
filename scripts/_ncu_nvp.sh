timeout 600 ncu --set full --import-source on --clock-control none -k regex:nvp_fwd_kernel -s 3 -c 1 -f -o gpurun_out/r5_nvp_fwd python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-micro > gpurun_out/r5_ncu.log 2>&1
timeout 600 ncu --set full --import-source on --clock-control none -k regex:nvp_pack_fwd_kernel -s 3 -c 1 -f -o gpurun_out/r5_nvp_pack python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-micro >> gpurun_out/r5_ncu.log 2>&1
ls -la gpurun_out/r5_nvp_*.ncu-rep
