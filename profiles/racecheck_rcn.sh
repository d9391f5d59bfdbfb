# racecheck of the two RCN kernels after the zero pad behind q (runs on the GPU box)
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 --target-processes all \
  python -m pytest tests/test_gpu_zz_late_additions.py tests/test_gpu_full_size.py -x -q \
  -k "one_pass and False-3-1.0 or config3_rcn_batched" > gpurun_out/sanitize_racecheck_rcn.txt 2>&1
echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/sanitize_racecheck_rcn.txt | tail -4
