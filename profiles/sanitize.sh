# compute-sanitizer over the tests of the last session's kernels (runs on the GPU box)
mkdir -p gpurun_out
sel='packed_batch_tail or configuration_major or binary_difference or one_pass'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --target-processes all \
  python -m pytest tests/test_gpu_zz_late_additions.py tests/test_gpu_logical_pull.py -x -q -k "$sel" > gpurun_out/sanitize_memcheck.txt 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds" gpurun_out/sanitize_memcheck.txt | tail -8
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --target-processes all \
  python -m pytest tests/test_gpu_zz_late_additions.py tests/test_gpu_logical_pull.py -x -q -k "packed_batch_tail and 36 or configuration_major and 70 or one_pass and 3-1.0" > gpurun_out/sanitize_racecheck.txt 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitize_racecheck.txt | tail -8
