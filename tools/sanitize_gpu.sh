# gpurun -- bash tools/sanitize_gpu.sh : compute-sanitizer memcheck + racecheck (+ synccheck, initcheck) over the parity
# tests of the value kernels (canonical tets order 1/2, thread-per-entity kernels, generic), small meshes.
# Summaries -> gpurun_out/sanitizer_<tool>.txt (copied to profiles/ by hand).
mkdir -p gpurun_out
T="tests/test_gpu_parity.py::test_values_canonical_vs_oracle tests/test_gpu_parity.py::test_values_rows_kernel_vs_oracle tests/test_gpu_parity.py::test_canonical_midsize_vs_c_oracle"
for tool in memcheck racecheck synccheck initcheck; do
  extra=""
  [ "$tool" = memcheck ] && extra="--leak-check no"
  [ "$tool" = racecheck ] && extra="--racecheck-report all"
  [ "$tool" = initcheck ] && T="tests/test_gpu_parity.py::test_values_canonical_vs_oracle"
  echo "== compute-sanitizer --tool $tool" > gpurun_out/sanitizer_$tool.txt
  ( time timeout 900 compute-sanitizer --tool $tool $extra --print-limit 20 --target-processes all \
      python -m pytest $T -x -q -p no:cacheprovider ) >> gpurun_out/sanitizer_$tool.txt 2>&1
  echo "exit code $?" >> gpurun_out/sanitizer_$tool.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|passed|failed|exit code|real" gpurun_out/sanitizer_$tool.txt | tail -6
done
