# gpurun -- bash tools/whatif_gpu.sh <n> name...: per-kernel ncu durations of timing-only what-if variants (results are wrong)
n=$1; shift
for v in "$@"; do
  echo "== variant $v"
  lib=$PWD/sucem-fem_b200/lib/variants/lib_$v.so
  [ "$v" = base ] && lib=$PWD/sucem-fem_b200/lib/libsucem_b200.so
  SUCEM_B200_LIB=$lib timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:k_canon -s 6 -c 2 python bench.py --n $n --steps 2 --warmup 3 --no-e2e --no-cpu --no-check 2>&1 | grep -E "gpu__time|inst_executed|wavefronts|issue_active" | head
done
