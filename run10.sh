timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; tail -3 gpurun_out/bench_r1g.err; cat gpurun_out/bench_r1g.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 250 --csv --log-file gpurun_out/launches_r1g.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-encoder > gpurun_out/ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_pair_kernel -s 4 -c 1 -o gpurun_out/scan_pair_r1g python bench.py --steps 3 --warmup 3 --no-cpu --no-encoder > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out/*.ncu-rep
