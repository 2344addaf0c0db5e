"""Both sharded-search protocols on real NCCL ranks against the unsharded search of the same index (bit-identical D and I):
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/check_sharded_nccl.py
The index is synthetic (seeded, generated on the GPUs); rank 0 also builds the WHOLE index on its GPU as the reference."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from densephrases_b200 import IvfPqIndex
from densephrases_b200.sharded import ShardedIvfPq

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
N, NLIST, K = 40_000_000, 4096, 10
lens = bench.uniform_lens(N, NLIST)
sh = ShardedIvfPq(NLIST, rank=rank, world=world, device=local).build_synthetic(bench.opq_matrix(7), lens, 7)
g = torch.Generator(device="cuda").manual_seed(99)
ok = True
for n, nprobe in ((1024, 64), (515, 32), (64, 256), (96, 16)):          # 515: ragged slices (the last rank's slice is padded)
    x = 0.5 * torch.randn((n, 768), generator=g, device="cuda")
    dist.broadcast(x, 0)
    sh.nprobe = nprobe
    res = {}
    for qs in (False, True):
        sh.query_split = qs
        D, I = sh.search_device(x, K)
        res[qs] = (D.clone(), I.clone())
    same = torch.equal(res[False][0], res[True][0]) and torch.equal(res[False][1], res[True][1])
    if rank == 0:
        full = IvfPqIndex(NLIST, device=local)
        full.set_opq(bench.opq_matrix(7)); full.gen_centroids(7); full.gen_pq(7); full.set_lists_synthetic(lens, 7)
        full.nprobe = nprobe
        Df, If = full.search(x, K)
        ref_ok = torch.equal(Df, res[True][0]) and torch.equal(If, res[True][1])
        del full
        torch.cuda.empty_cache()
        print(f"n={n} nprobe={nprobe} world={world}: query-split == list-split: {same}; == unsharded index: {ref_ok}", flush=True)
        ok = ok and same and ref_ok
    dist.barrier()
if rank == 0:
    print("SHARDED NCCL CHECK", "PASSED" if ok else "FAILED", flush=True)
dist.destroy_process_group()
