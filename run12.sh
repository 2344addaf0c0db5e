timeout 600 python -m pytest tests/test_search_gpu.py -x -q 2>&1 | tail -12
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 --scale 0.1 2>&1 | tail -30 | cut -c1-400
