export PYTHONUNBUFFERED=1
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 8 python -m pytest "tests/test_search_gpu.py::test_search_matches_oracle" -q -x -k "16-5000 or 8-100" 2>&1 | tail -12
echo racecheck_rc=$?
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest "tests/test_encoder.py::test_cuda_encoder_matches_reference_fixture" "tests/test_gemm_gpu.py" -q -x -k "b3_s24 or 128-128-32 or 100-256-64" 2>&1 | tail -8
echo memcheck_enc_rc=$?
