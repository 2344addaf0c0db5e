set -x
free -g | head -2
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r1a.json 2> gpurun_out/bench_r1a.err; tail -3 gpurun_out/bench_r1a.err; cat gpurun_out/bench_r1a.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_r1a.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_kernelILi0 -s 4 -c 2 -o gpurun_out/scan_r1a python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out
