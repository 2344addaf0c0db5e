timeout 900 python -m pytest tests/test_search_gpu.py -x -q 2>&1 | tail -15
timeout 900 python bench.py --steps 20 --warmup 3 --no-encoder > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; tail -3 gpurun_out/bench_r1f.err; python -c "
import json; d=json.load(open('gpurun_out/bench_r1f.json')); print(d['value'], d['e2e']['value'], d['roofline'], d['cpu_baseline'], d['exact_fallback_queries_last_batch'])"
