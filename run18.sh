timeout 900 python -m pytest tests/test_search_gpu.py -q -x 2>&1 | tail -12
python tools/bench_c4_shard.py 32 256 2>&1 | tail -2
timeout 600 python bench.py --steps 50 --no-cpu --no-encoder 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('N1', d['value'], d['e2e']['value'], d['ms_per_step'], d['roofline']['kernel_ms'])"
