# gpurun -- bash tools/variants_gpu.sh <n> name1 name2 ...: step time and per-kernel ncu durations of library variants
n=$1; shift
for v in "$@"; do
  echo "== variant $v"
  lib=$PWD/sucem-fem_b200/lib/variants/lib_$v.so
  [ "$v" = base ] && lib=$PWD/sucem-fem_b200/lib/libsucem_b200.so
  SUCEM_B200_LIB=$lib timeout 300 python bench.py --n $n --steps 10 --no-e2e --no-cpu 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('n$n', d['value'], d['ms_per_step'], d['roofline']['frac'], d['setup_ms'], d.get('checked'))"
  SUCEM_B200_LIB=$lib timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_canon -s 6 -c 2 python bench.py --n $n --steps 2 --warmup 3 --no-e2e --no-cpu --no-check 2>&1 | grep -E "gpu__time" | head
done
