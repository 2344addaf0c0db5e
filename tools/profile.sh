#!/usr/bin/env bash
# One-GPU profiling pass (run on the GPU box: gpurun --timeout 2400 -- 'bash tools/profile.sh TAG').  Never wraps a multi-rank command.
# Writes into gpurun_out/: per-launch lists (one C2 step, one C4 rank step at nprobe 32 and 256, one encoder forward per mode) and
# `ncu --set full` captures of the dominant kernels; then profiles/traffic.json (tools/make_traffic.py, keyed by the csrc hash) and the
# text summaries (tools/ncu_summary.py).  Copy what should be judged from gpurun_out/ into profiles/.
set -u
TAG=${1:-r2}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --clock-control none"
B="python bench.py --steps 3 --warmup 3 --no-cpu --no-encoder --no-c1 --no-c4"

timeout 300 $NCU --metrics gpu__time_duration.sum -c 3000 --csv --log-file $OUT/${TAG}_launches_c2.csv $B > /dev/null 2>&1
for np in 32 256; do
  timeout 400 $NCU --metrics gpu__time_duration.sum -c 4000 --csv --log-file $OUT/${TAG}_launches_shard_c4_np${np}.csv python tools/bench_shard.py c4 $np > /dev/null 2>&1
done
for mode in tf32 bf16x3; do
  timeout 300 $NCU --metrics gpu__time_duration.sum -c 6000 --csv --log-file $OUT/${TAG}_launches_encoder_${mode}.csv python tools/bench_encoder.py 64 64 $mode 2 > /dev/null 2>&1
done

timeout 600 $NCU --set full --import-source on -k regex:scan_quad_kernel -s 4 -c 1 -o $OUT/${TAG}_scan_quad_c2 $B > $OUT/ncu1.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:scan_quad_kernel -s 12 -c 1 -o $OUT/${TAG}_scan_quad_c4shard python tools/bench_shard.py c4 256 > $OUT/ncu2.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:^scan_kernel$ -s 12 -c 1 -o $OUT/${TAG}_scan_single_c4shard python tools/bench_shard.py c4 32 > $OUT/ncu3.log 2>&1
# QKV projection of the second forward (48 GEMM launches per forward) and one attention launch, bf16x3 mode
timeout 600 $NCU --set full --import-source on -k regex:gemm_bf16x3_persist_kernel -s 48 -c 1 -o $OUT/${TAG}_gemm_bf16x3 python tools/bench_encoder.py 64 64 bf16x3 2 > $OUT/ncu4.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:attention_tc_bx_kernel -s 12 -c 1 -o $OUT/${TAG}_attention_tc_bx python tools/bench_encoder.py 64 64 bf16x3 2 > $OUT/ncu5.log 2>&1
timeout 600 $NCU --set full --import-source on -k regex:gemm_tf32_persist_kernel -s 48 -c 1 -o $OUT/${TAG}_gemm_tf32_persist python tools/bench_encoder.py 64 64 tf32 2 > $OUT/ncu6.log 2>&1

ARGS=""
[ -f $OUT/${TAG}_scan_quad_c2.ncu-rep ] && ARGS="$ARGS scan_quad_kernel|C2|nprobe256|n1=$OUT/${TAG}_scan_quad_c2.ncu-rep"
[ -f $OUT/${TAG}_scan_quad_c4shard.ncu-rep ] && ARGS="$ARGS scan_quad_kernel|C4|nprobe256|n8=$OUT/${TAG}_scan_quad_c4shard.ncu-rep"
[ -f $OUT/${TAG}_scan_single_c4shard.ncu-rep ] && ARGS="$ARGS scan_kernel<FAST>|C4|nprobe32|n8=$OUT/${TAG}_scan_single_c4shard.ncu-rep"
python tools/make_traffic.py $ARGS
cp profiles/traffic.json $OUT/${TAG}_traffic.json
for rep in scan_quad_c2 scan_quad_c4shard scan_single_c4shard gemm_bf16x3 attention_tc_bx gemm_tf32_persist; do
  [ -f $OUT/${TAG}_${rep}.ncu-rep ] && python tools/ncu_summary.py $OUT/${TAG}_${rep}.ncu-rep $OUT/${TAG}_${rep}_ncu_summary.txt > /dev/null
done
ls -la $OUT | tail -30
