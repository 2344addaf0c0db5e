export PYTHONUNBUFFERED=1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python -m pytest "tests/test_search_gpu.py::test_search_matches_oracle" "tests/test_search_gpu.py::test_ragged_empty_lists_and_explicit_ids" "tests/test_search_gpu.py::test_list_range_shards_on_one_device" "tests/test_search_gpu.py::test_coarse_ties_duplicate_centroids" "tests/test_mips.py::test_fused_window_scores_match_reconstruct_path" -q -x 2>&1 | tail -15
echo memcheck_rc=$?
