python tools/bench_sgemm.py
timeout 600 python -m pytest tests/test_search_gpu.py -x -q 2>&1 | tail -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 50 --warmup 3 > gpurun_out/bench_r1h_n2.json 2> gpurun_out/bench_r1h_n2.err; tail -3 gpurun_out/bench_r1h_n2.err; cat gpurun_out/bench_r1h_n2.json
