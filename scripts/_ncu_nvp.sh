timeout 600 ncu --set full --import-source on --clock-control none -k regex:nvp_bwd_kernel -s 3 -c 1 -f -o gpurun_out/r3c_nvp_bwd python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-graph --no-micro > gpurun_out/r3c_ncu.log 2>&1
ls -la gpurun_out/r3c_nvp_bwd.ncu-rep
