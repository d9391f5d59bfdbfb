# compute-sanitizer memcheck over the WHOLE GPU test suite (runs on the GPU box)
mkdir -p gpurun_out
timeout 560 compute-sanitizer --tool memcheck --error-exitcode 9 \
  python -m pytest tests -m gpu -q -x -p no:cacheprovider > gpurun_out/sanitize_memcheck_all.txt 2>&1
echo "memcheck rc=$?"; grep -E "ERROR SUMMARY|passed|failed|Invalid|out of bounds|Error" gpurun_out/sanitize_memcheck_all.txt | tail -8
