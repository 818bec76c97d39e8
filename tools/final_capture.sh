# Round-end evidence for profiles/ (gpurun -- bash tools/final_capture.sh): the ncu launch list of the bench command and one
# --set full capture of the kernels of a step.  (Bench lines themselves come from un-profiled runs.)
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras --no-check > /dev/null 2>&1
grep -c k_canon gpurun_out/launches.csv
timeout 400 ncu --set full --clock-control none --import-source on -k regex:"k_canon|k_pack" -s 9 -c 3 -f -o gpurun_out/prof_canon_n119 \
  python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --no-extras --no-check > /dev/null 2>&1
ls -la gpurun_out/
