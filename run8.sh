timeout 900 ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k regex:scan_kernelILi0 -s 4 -c 1 -o gpurun_out/scan_r1c python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep
