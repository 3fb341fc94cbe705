timeout 600 ncu --set full --import-source on --clock-control none -k regex:pdf_merge -c 1 -f -o gpurun_out/r3l_pdf python scripts/micro_hbm.py --once > gpurun_out/r3l_ncu_pdf.log 2>&1
ls -la gpurun_out/r3l_pdf.ncu-rep
