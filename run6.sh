timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_r1d.json 2> gpurun_out/bench_r1d.err; tail -3 gpurun_out/bench_r1d.err; cat gpurun_out/bench_r1d.json
timeout 900 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_r1d_ref.json 2> gpurun_out/bench_r1d_ref.err; tail -3 gpurun_out/bench_r1d_ref.err; cat gpurun_out/bench_r1d_ref.json
