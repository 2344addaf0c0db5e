#!/usr/bin/env bash
# One-GPU profiling pass (run on the GPU box: gpurun --timeout 1500 -- 'bash tools/profile.sh TAG').  Never wraps a multi-rank command.
# Writes into gpurun_out/: bench line, per-launch list of a short bench run, `ncu --set full` captures of the dominant kernels
# (search: scan_pair_kernel; encoder: gemm_tf32_persist_kernel, attention_tc_kernel) and their text summaries.
# Copy what should be judged into profiles/ (tools/ncu_summary.py output is what profiles/*_ncu_summary.txt hold).
set -u
TAG=${1:-rX}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"

timeout 900 python bench.py > $OUT/bench_${TAG}_n1.json 2> $OUT/bench_${TAG}_n1.err
tail -2 $OUT/bench_${TAG}_n1.err; cat $OUT/bench_${TAG}_n1.json

# per-launch list of a 3-step search bench (no CPU leg, no encoder) and of one encoder forward
timeout 600 $NCU --metrics gpu__time_duration.sum -c 400 --csv --log-file $OUT/launches_${TAG}_bench_steps3.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-encoder > $OUT/ncu_launches.log 2>&1
timeout 300 $NCU --metrics gpu__time_duration.sum -k regex:"gemm_tf32|attention|layernorm|embed_ln|gather_cls" -s 87 -c 87 --csv \
    --log-file $OUT/launches_${TAG}_encoder.csv python tools/bench_encoder.py > $OUT/ncu_enc_launches.log 2>&1

# full captures: one launch each, after warm-up launches
timeout 900 $NCU --set full --import-source on -k regex:scan_pair_kernel -s 4 -c 1 -o $OUT/scan_pair_${TAG} \
    python bench.py --steps 3 --warmup 3 --no-cpu --no-encoder > $OUT/ncu_scan.log 2>&1
# the QKV projection (first GEMM of a layer; skip the first forward = 48 GEMM launches) and one attention launch
timeout 600 $NCU --set full --import-source on -k regex:gemm_tf32_persist_kernel -s 48 -c 1 -o $OUT/gemm_persist_${TAG} \
    python tools/bench_encoder.py > $OUT/ncu_gemm.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:attention_tc_kernel -s 12 -c 1 -o $OUT/attention_tc_${TAG} \
    python tools/bench_encoder.py > $OUT/ncu_attn.log 2>&1

for rep in scan_pair gemm_persist attention_tc; do
    [ -f $OUT/${rep}_${TAG}.ncu-rep ] && python tools/ncu_summary.py $OUT/${rep}_${TAG}.ncu-rep $OUT/${TAG}_${rep}_ncu_summary.txt
done
ls -la $OUT | tail -20
