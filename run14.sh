timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/enc_launches.csv python tools/bench_encoder.py 64 64 > gpurun_out/enc_b.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tf32_kernel -s 40 -c 2 -o gpurun_out/gemm_r1 python tools/bench_encoder.py 64 64 > gpurun_out/enc_c.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
