# gpurun -- bash tools/edge2_gpu.sh : parity of the lane-per-item edge kernel + A/B timing against the lane-pair kernel
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_api_gpu.py -x -q 2>&1 | tail -15
for ek in 1 2; do
  for n in 70 119; do
    SFEM_EDGE_KERNEL=$ek timeout 300 python bench.py --n $n --steps 10 --no-e2e --no-cpu 2>gpurun_out/edge2_err_$ek_$n.txt | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('edge kernel $ek n$n', d['value'], d['ms_per_step'], d['roofline']['frac'], d['setup_ms'], d.get('checked'))"
  done
  SFEM_EDGE_KERNEL=$ek timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_canon -s 6 -c 2 python bench.py --n 119 --steps 2 --warmup 3 --no-e2e --no-cpu --no-check 2>&1 | grep -E "k_canon|gpu__time" | head
done
