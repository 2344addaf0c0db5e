#!/usr/bin/env python
"""C5 (BASELINE.json configs[4]): the evaluation loop of eval_phrase_retrieval.py over a multi_wiki-scale synthetic dump, one
process per GPU:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29544 \
        tools/run_c5_eval.py [--N 580000000 --nlist 1048576 --questions 1280 --top_k 10,40]

Every rank runs the SAME driver code the reference runs in one process (load_encoder -> embed_all_query -> load_phrase_index ->
mips.search per batch of 64 questions -> metrics): the encoder is replicated, the phrase index is sharded by list range
(MIPS builds a ShardedIvfPq when WORLD_SIZE > 1) and every `mips.search` is a collective.  When the unmodified reference script
is present (/root/reference/eval_phrase_retrieval.py; build container only) its own `evaluate` is the loop that runs; on the GPU
box, where the reference tree does not exist, densephrases_b200.runtime.evaluate (the restatement pinned to it by
tests/test_api.py) runs instead.  Synthetic dump = spec files (densephrases_b200/synthetic_dump.py); encoder = seeded random
weights (DPH_ALLOW_RANDOM_INIT=1).  Prints one JSON line on rank 0: questions/s of the search loop and the stage breakdown."""
import argparse
import importlib.util
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--N", type=int, default=580_000_000)
    ap.add_argument("--nlist", type=int, default=1_048_576)
    ap.add_argument("--questions", type=int, default=1280)
    ap.add_argument("--top_k", default="10,40")
    ap.add_argument("--tokens_per_doc", type=int, default=128)
    a = ap.parse_args()
    real_stdout = os.dup(1)            # NCCL prints its version banner on stdout: keep fd 1 for the one JSON line
    os.dup2(2, 1)
    import numpy as np
    import torch
    os.environ["DPH_ALLOW_RANDOM_INIT"] = "1"
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    from densephrases import Options
    from densephrases_b200 import runtime as R
    from densephrases_b200 import synthetic_dump as SD
    from densephrases_b200.mips import distributed_context
    distributed_context()
    work = tempfile.mkdtemp(prefix=f"dph_c5_r{rank}_")
    index_name = f"start/{a.nlist}_flat_OPQ96"                      # the reference's naming (build_phrase_index.py:24-25); 'PQ' switches PQ mode
    ntotal = SD.write_synthetic_dump(work, index_name, a.N, a.nlist, a.tokens_per_doc)
    qa = SD.write_synthetic_questions(os.path.join(work, "questions.json"), a.questions)
    os.makedirs(os.path.join(work, "ckpt"), exist_ok=True)
    evaluate_fn, which = R.evaluate, "densephrases_b200.runtime.evaluate (restatement of eval_phrase_retrieval.py:49-91)"
    ref = "/root/reference/eval_phrase_retrieval.py"
    if os.path.exists(ref):
        spec = importlib.util.spec_from_file_location("ref_eval_phrase_retrieval", ref)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        evaluate_fn, which = mod.evaluate, "UNMODIFIED /root/reference/eval_phrase_retrieval.py:evaluate"
    o = Options()
    o.add_model_options(); o.add_index_options(); o.add_retrieval_options(); o.add_data_options()
    base = ["--run_mode", "eval", "--cuda", "--dump_dir", work, "--index_name", index_name, "--load_dir", os.path.join(work, "ckpt"),
            "--test_path", qa, "--eval_batch_size", "64", "--aggregate"]
    out = {"config": {"workload": f"C5: {ntotal}-phrase IVF{a.nlist},PQ96 (OPQ96) synthetic dump, {a.questions} questions, eval batch 64, nprobe 256",
                      "n_gpus": world, "loop": which}, "runs": []}
    args = o.parse(base + ["--top_k", "10"])
    t0 = time.time()
    enc, tok, _ = R.load_encoder("cuda", args)
    t_enc_load = time.time() - t0
    t0 = time.time()
    mips = R.load_phrase_index(args)
    torch.cuda.synchronize()
    t_index = time.time() - t0
    # one untimed batch: first-call allocations (probe / LUT / candidate workspaces, the split centroid copy of the tensor-core coarse quantizer)
    t0 = time.time()
    rng = np.random.default_rng(0)
    mips.search(rng.standard_normal((64, 1536)), q_texts=["warm-up"] * 64, top_k=10, aggregate=True)
    torch.cuda.synchronize()
    out["load_seconds"] = {"encoder": t_enc_load, "index": t_index, "first_search_batch": time.time() - t0}
    for top_k in [int(v) for v in a.top_k.split(",")]:
        args = o.parse(base + ["--top_k", str(top_k)])
        for k_ in mips.stage_seconds:
            mips.stage_seconds[k_] = 0 if k_ == "batches" else 0.0
        # encoder stage timed on its own (the loop below repeats it inside evaluate, as the reference does)
        _, questions, _, _ = R.load_qa_pairs(args.test_path, args)
        torch.cuda.synchronize(); t0 = time.time()
        R.embed_all_query(questions, args, enc, tok)
        torch.cuda.synchronize(); t_embed = time.time() - t0
        for k_ in mips.stage_seconds:
            mips.stage_seconds[k_] = 0 if k_ == "batches" else 0.0
        t0 = time.time()
        res = evaluate_fn(args, mips, enc, tok)
        torch.cuda.synchronize()
        t_total = time.time() - t0
        st = dict(mips.stage_seconds)
        t_search = st["mips"] + st["get_idxs"] + st["phrase_vectors"] + st["phrase_select"] + st["metadata"]
        out["runs"].append({"top_k": top_k, "questions": len(questions), "evaluate_seconds": t_total, "embed_all_query_seconds": t_embed,
                            "search_loop_seconds": t_search, "questions_per_s_search_loop": len(questions) / t_search,
                            "questions_per_s_encoder": len(questions) / t_embed,
                            "questions_per_s_end_to_end": len(questions) / (t_embed + t_search),
                            "stage_seconds": st, "vector_queries_per_s_index_only": 2 * len(questions) / st["mips"]})
    if rank == 0:
        os.write(real_stdout, (json.dumps(out) + "\n").encode())
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
