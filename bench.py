#!/usr/bin/env python
"""bench.py -- queries/sec top-10 over the synthetic PQ96 phrase index (BASELINE.json metric), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
  python bench.py --impl reference [...]                          # the reference's CPU path (FAISS-equivalent restatement)

A *step* is one pass of the hot path (OPQ rotation -> coarse top-nprobe -> LUT -> PQ96 scan -> top-k merge) over one
batch of synthetic d=768 vector queries.  N=1 workload = BASELINE.json configs[1] (C2): 100M phrases, IVF4096,PQ96,
batch 64, nprobe 256 (the reference's fixed value, densephrases/index.py:53,62), k=10.  N>1: weak scaling -- N x 100M
phrases, IVF(4096 N), list-range shards, batch 64 N, one all-gather of per-shard top-k.
Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED_INDEX, SEED_QUERY = 1234, 4321
D, K_TOP, NPROBE = 768, 10, 256
PER_GPU_N, PER_GPU_NLIST, PER_GPU_BATCH = 100_000_000, 4096, 64


def opq_matrix(seed):
    rng = np.random.default_rng(seed)
    return np.linalg.qr(rng.standard_normal((D, D)))[0].astype(np.float32)


def uniform_lens(N, nlist):
    base, rem = divmod(N, nlist)
    lens = np.full(nlist, base, dtype=np.int64)
    lens[:rem] += 1
    return lens


def workload(n_gpus, scale=1.0):
    N = int(PER_GPU_N * scale) * n_gpus
    nlist = PER_GPU_NLIST * n_gpus
    return dict(N=N, nlist=nlist, batch=PER_GPU_BATCH * n_gpus, nprobe=NPROBE, k=K_TOP, lens=uniform_lens(N, nlist))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    Q = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu):
        self.gpu, self.p = gpu, None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20", "-i", str(self.gpu)],
                                      stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self, t0=None, t1=None):
        """Summary over samples whose timestamp falls inside [t0, t1] (time.time() seconds) when given."""
        import datetime
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            out = self.p.communicate(timeout=5)[0]
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 8:
                continue
            try:
                ts = datetime.datetime.strptime(f[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                if t0 is not None and not (t0 - 0.02 <= ts <= t1 + 0.02):
                    continue
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v == "Active":
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "samples": len(sm),
                "reasons": sorted(reasons)}


def make_queries(ix, wl, nbatches, rank, world, device):
    """SURVEY 8d 'near' queries: q = A^T (centroid + decode(code_j)) + N(0, 0.3^2) for random stored j; same on all ranks."""
    import torch
    g = torch.Generator().manual_seed(SEED_QUERY)
    total = nbatches * wl["batch"]
    ids = torch.randint(0, wl["N"], (total,), generator=g, dtype=torch.int64)
    noise = torch.randn((total, D), generator=g, dtype=torch.float32) * 0.3
    v, _ = ix.local.reconstruct_batch(ids.to(device))
    if world > 1:
        import torch.distributed as dist
        dist.all_reduce(v)
    A = torch.from_numpy(ix.local.opq_matrix()).to(device)
    q = v @ A + noise.to(device)
    return q.reshape(nbatches, wl["batch"], D).contiguous()


def ref_index_for(wl, oracle):
    A = opq_matrix(SEED_INDEX)
    return oracle.RefIndex(A, oracle.gen_pq(SEED_INDEX), wl["lens"], centroids=oracle.gen_centroids(SEED_INDEX, 0, wl["nlist"]), seed=SEED_INDEX)


def cpu_time_queries(ref, x, k, nprobe, oracle, resident_budget_gb=24.0):
    """Time the oracle (FAISS-equivalent CPU restatement, OpenMP over queries like faiss parallel_mode 0) on queries x.
    Lists probed by the sample are materialised in RAM first (faiss scans resident inverted lists); untimed."""
    xr = ref.rotate(x)
    _, key = ref.coarse(xr, nprobe)
    lists = np.unique(key[key >= 0])
    need_gb = float(ref.list_len[lists].sum()) * 96 / 1e9
    kind_note = "resident lists"
    rr = ref
    if need_gb <= resident_budget_gb:
        rr = ref.with_resident_lists(lists)
    else:
        kind_note = "lists regenerated on the fly (RAM budget)"
    t0 = time.perf_counter()
    xr = rr.rotate(x)
    _, key = rr.coarse(xr, nprobe)
    Dr, Ir = rr.search_preassigned(xr, key, k)
    dt = time.perf_counter() - t0
    return dt, Dr, Ir, kind_note


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from oracle import ivfpq_ref as oracle
    oracle.build()
    wl = workload(args.gpus, args.scale)
    ref = ref_index_for(wl, oracle)
    cores = oracle.lib().ref_num_threads()
    nq = min(wl["batch"], 64)
    rng = np.random.default_rng(SEED_QUERY)
    ids = rng.integers(0, wl["N"], nq)
    v, _ = ref.reconstruct(ids)
    x = (v @ ref.A + 0.3 * rng.standard_normal((nq, D))).astype(np.float32)
    xr = ref.rotate(x)
    _, key = ref.coarse(xr, wl["nprobe"])
    lists = np.unique(key[key >= 0])
    need_gb = float(ref.list_len[lists].sum()) * 96 / 1e9
    rr = ref.with_resident_lists(lists) if need_gb <= 40.0 else ref
    times = []
    for s in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        xr = rr.rotate(x)
        _, key = rr.coarse(xr, wl["nprobe"])
        rr.search_preassigned(xr, key, wl["k"])
        times.append(time.perf_counter() - t0)
    t = sum(times[args.warmup:])
    qps = nq * args.steps / t
    sample = f"{nq} queries/step over the probed lists ({need_gb:.1f} GB, {'resident' if rr is not ref else 'regenerated'}) of the {wl['N']}-phrase index"
    line = {"impl": "reference", "metric": "queries/sec top-10 over PQ96 phrase index", "value": qps, "unit": "queries/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * t / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_dict(wl, args.gpus, "cpu"),
            "cpu_baseline": {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": qps, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    emit(line)
    return 0


def config_dict(wl, n_gpus, where):
    return {"workload": f"C2 x{n_gpus}: {wl['N']}-phrase IVF{wl['nlist']},PQ96 (OPQ96) index, batch {wl['batch']} d=768 near queries, "
                        f"nprobe {wl['nprobe']}, top-{wl['k']}", "N": wl["N"], "nlist": wl["nlist"], "batch": wl["batch"], "nprobe": wl["nprobe"],
            "k": wl["k"], "parallelism": f"list-range shards x{n_gpus}" if n_gpus > 1 else "1 gpu", "where": where,
            "l2": "index (9.6 GB/GPU) is larger than L2; every step uses a different query batch"}


def run_ours(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    from densephrases_b200.sharded import ShardedIvfPq
    wl = workload(world, args.scale)
    W, K = args.warmup, args.steps
    ix = ShardedIvfPq(wl["nlist"], rank=rank, world=world, device=local_rank)
    ix.build_synthetic(opq_matrix(SEED_INDEX), wl["lens"], SEED_INDEX)
    ix.nprobe = wl["nprobe"]
    torch.cuda.synchronize()
    nb = W + K
    Q = make_queries(ix, wl, nb, rank, world, dev)
    Qh = Q.cpu().pin_memory()
    k = wl["k"]

    def barrier():
        if world > 1:
            import torch.distributed as dist
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        import torch.distributed as dist
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def stage(msg):
        if rank == 0 and args.verbose:
            print(f"[bench] {msg}", file=sys.stderr, flush=True)

    stage("index built, queries made")
    # ---- device-resident timing (value) ----
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for s in range(W):
        ix.search_device(Q[s], k)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_wall0 = time.time()
    e0.record()
    for s in range(W, W + K):
        Dd, Id = ix.search_device(Q[s], k)
    e1.record()
    barrier()
    t_wall1 = time.time()
    ms_dev = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop(t_wall0, t_wall1) if rank == 0 else None
    last_dev = (Dd.cpu().numpy(), Id.cpu().numpy())
    stage(f"device pass done: {ms_dev / K:.3f} ms/step")

    # ---- end to end through the host API (pinned host in, host out) ----
    for s in range(W):
        ix.search(Qh[s], k)
    barrier()
    e0.record()
    for s in range(W, W + K):
        Dh, Ih = ix.search(Qh[s] if world > 1 else Qh[s].numpy(), k)
    e1.record()
    barrier()
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    assert np.array_equal(np.asarray(Dh), last_dev[0]) and np.array_equal(np.asarray(Ih), last_dev[1])
    stage(f"e2e pass done: {ms_e2e / K:.3f} ms/step")

    # ---- roofline of the dominant kernel (PQ scan), CUDA events around the kernel itself ----
    # (a) algorithmic bytes of each timed batch (untimed pass), (b) the same K steps back to back with events around the scan
    alg_bytes = []
    lens = wl["lens"]
    lo, hi = ix.range
    for s in range(W, W + K):
        ix.local.search_partial(Q[s], k) if world > 1 else ix.local.search(Q[s], k)
        pr = ix.local.last_probes(wl["batch"]).astype(np.int64)
        m = (pr >= lo) & (pr < hi)
        alg_bytes.append(float(lens[pr[m]].sum()) * 96.0)
    ix.local.set_profile(True)
    barrier()
    e0.record()
    for s in range(W, W + K):
        ix.local.search_partial(Q[s], k) if world > 1 else ix.local.search(Q[s], k)
    e1.record()
    torch.cuda.synchronize()
    ms_prof_pass = e0.elapsed_time(e1)
    scan_ms = [float(v) for v in ix.local.profile_scan_ms()][-K:]
    ix.local.set_profile(False)
    flags = int(ix.local.last_flags(wl["batch"]).sum())
    pk, pk_kind = peaks()
    t_scan = sum(scan_ms) / len(scan_ms) / 1000.0
    achieved = (sum(alg_bytes) / len(alg_bytes)) / t_scan / 1e9
    pair_mode = ix.local.last_used_pair_mode()
    traffic = None
    tp = os.path.join(ROOT, "profiles", "scan_traffic.json")
    limiter = None
    if os.path.exists(tp):
        tj = json.load(open(tp))
        traffic = tj.get("dram_bytes_per_launch") if pair_mode else tj.get("single_query_kernel", {}).get("dram_bytes_per_launch")
        if pair_mode and tj.get("lsu_data_pipe_pct"):      # ncu: the pair kernel's own limiter is the LSU data pipe, not DRAM
            limiter = {"pipe": "lsu data pipe (shared-memory gathers + code loads)", "pct_of_peak": tj["lsu_data_pipe_pct"], "source": tj.get("source")}
    roofline = {"kernel": "scan_pair_kernel" if pair_mode else "scan_kernel<FAST>", "gathers": "two queries per gather (pair-packed)" if pair_mode else "one query per gather", "bound": "hbm", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                "peak_kind": pk_kind, "traffic": traffic, "kernel_ms": 1000.0 * t_scan, "algorithmic_bytes_per_launch": sum(alg_bytes) / len(alg_bytes),
                "share_of_step": 1000.0 * t_scan / (ms_prof_pass / K), "step_ms_same_pass": ms_prof_pass / K,
                "how": "CUDA events around the kernel inside a back-to-back K-step loop on the launching stream", "limiter": limiter}

    # ---- CPU baseline beside it (rank 0, N=1 only): the oracle on a bounded sample of the same workload ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        from oracle import ivfpq_ref as oracle
        oracle.build()
        ref = ref_index_for(wl, oracle)
        nq = min(wl["batch"], 64)
        x = Qh[W + K - 1][:nq].numpy()
        dt, Dr, Ir, note = cpu_time_queries(ref, x, k, wl["nprobe"], oracle)
        same = bool(np.array_equal(Dr.view(np.int32), last_dev[0][:nq].view(np.int32)) and np.array_equal(Ir, last_dev[1][:nq]))
        cpu = {"value": nq / dt, "unit": "queries/s", "cores": oracle.lib().ref_num_threads(), "kind": "port",
               "sample": f"{nq} queries of the last timed batch, full nprobe={wl['nprobe']} scan, {note}; GPU results bit-identical: {same}"}

    # ---- C3: query encoder (2 x SpanBERT-base towers, random-init weights, synthetic tokens) + search, device resident ----
    enc_info = None
    if rank == 0 and world == 1 and not args.no_encoder:
        from densephrases_b200.encoder import BertGeometry, Encoder, random_state_dict, synthetic_query_batch
        geo = BertGeometry()
        enc = Encoder(geo, state_dict=random_state_dict(geo, 1), device=local_rank)
        ids, mask, tt = (t.to(dev) for t in synthetic_query_batch(64, 64, geo.vocab_size, 2))
        enc_info = {}
        for mode in ("tf32", "3xtf32"):
            enc.set_precision(mode == "3xtf32")
            for _ in range(3):
                enc.embed_query(ids, mask, tt)
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                qs, qe = enc.embed_query(ids, mask, tt)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            flops = 64 * 22.05e9
            enc_info[mode] = {"ms_per_64_questions": ms, "questions_per_s": 64000.0 / ms, "tensor_tflops": flops / ms / 1e9,
                              "mma_tflops_issued": flops * (3.0 if mode == "3xtf32" else 1.0) / ms / 1e9,
                              "tensor_pipe_frac_of_tf32_peak": flops * (3.0 if mode == "3xtf32" else 1.0) / ms / 1e9 / (pk["bf16_tflops"] / 2.0)}
        # end to end: 64 questions -> encoder (tf32) -> ONE stacked [128,768] search (start rows then end rows, index.py:195-202)
        enc.set_precision(False)
        for _ in range(2):
            qs, qe = enc.embed_query(ids, mask, tt)
            ix.search_device(torch.cat([qs[:, 0], qe[:, 0]], 0).contiguous(), k)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(5):
            qs, qe = enc.embed_query(ids, mask, tt)
            ix.search_device(torch.cat([qs[:, 0], qe[:, 0]], 0).contiguous(), k)
        e1.record()
        torch.cuda.synchronize()
        enc_info["c3_questions_per_s"] = 64 * 5 / (e0.elapsed_time(e1) / 1000.0)
        enc_info["note"] = "B=64,S=64; 22.05 GFLOP/question over both towers; GEMMs on tcgen05 kind::tf32 (peak taken as half the measured bf16 peak)"
        del enc

    if rank == 0:
        B = wl["batch"]
        line = {"metric": "queries/sec top-10 over PQ96 phrase index", "value": B * K / (ms_dev / 1000.0), "unit": "queries/s", "n_gpus": world,
                "steps": K, "warmup": W, "ms_per_step": ms_dev / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic", "config": config_dict(wl, world, "hbm"),
                "e2e": {"value": B * K / (ms_e2e / 1000.0), "unit": "queries/s", "h2d_bytes_per_step": B * D * 4, "d2h_bytes_per_step": B * k * 12,
                        "ms_per_step": ms_e2e / K},
                "gpu_launches": K * ((18 if pair_mode else 12) + (3 if world > 1 else 0)), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
                "exact_fallback_queries_last_batch": flags, "encoder": enc_info}
        emit(line)
    del ix
    if world > 1:
        import torch.distributed as dist
        dist.destroy_process_group()
    return 0


_REAL_STDOUT = None


def quiet_stdout():
    """The contract is ONE JSON line on stdout; NCCL prints a version banner there (seen on the GPU box).  Route fd 1 to stderr for
    the whole run and keep the real stdout for the final line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None, help="default: 100 (ours), 10 (--impl reference: each step is ~1.5 s of all-core CPU work)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--scale", type=float, default=1.0, help="debug only: shrink the per-GPU index (the headline run uses 1.0)")
    ap.add_argument("--verbose", action="store_true")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-encoder", action="store_true", help="skip the C3 encoder leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.steps is None:
        args.steps = 10 if args.impl == "reference" else 100
    quiet_stdout()
    return run_reference(args) if args.impl == "reference" else run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
