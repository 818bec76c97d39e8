for v in "$@"; do
  echo "== variant $v"
  SUCEM_B200_LIB=$PWD/sucem-fem_b200/variants/lib_$v.so timeout 300 python bench.py --n 70 --steps 5 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('n70', d['value'], d['ms_per_step'], d['roofline']['frac'])"
  SUCEM_B200_LIB=$PWD/sucem-fem_b200/variants/lib_$v.so ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_canon -s 6 -c 2 python bench.py --n 70 --steps 2 --warmup 3 --no-e2e --no-cpu 2>&1 | grep -E "gpu__time" | head
done
