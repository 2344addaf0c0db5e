timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/bench_r1c.json 2> gpurun_out/bench_r1c.err; tail -3 gpurun_out/bench_r1c.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1c.json')); print(d['value'], d['e2e']['value'], d['roofline'])"
