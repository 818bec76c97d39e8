# Round-end evidence for profiles/ (gpurun -- bash tools/final_capture.sh): un-profiled bench lines (both arms), the ncu
# launch list of the bench command and one --set full capture of the two kernels of a step.
mkdir -p gpurun_out
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_line.json 2> gpurun_out/bench_line.err
tail -c 600 gpurun_out/bench_line.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
tail -c 300 gpurun_out/bench_reference.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
grep -c k_canon gpurun_out/launches.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_canon -s 6 -c 2 -f -o gpurun_out/prof_canon_n119 \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu > /dev/null 2>&1
ls -la gpurun_out/
