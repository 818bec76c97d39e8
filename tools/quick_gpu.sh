# parity of the value kernels + the two bench sizes without the e2e / CPU legs (gpurun -- bash tools/quick_gpu.sh)
timeout 300 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -2
timeout 200 python bench.py --n 70 --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('n70', d['value'], d['ms_per_step'], d['roofline']['frac'], d['setup_ms'])"
timeout 200 python bench.py --steps 10 --no-e2e --no-cpu 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('n119', d['value'], d['ms_per_step'], d['roofline']['frac'], d['setup_ms'])"
