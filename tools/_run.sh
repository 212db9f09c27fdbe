python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/pytest56.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/pytest56.log
python bench.py --no-cpu-baseline > gpurun_out/bench56.json 2> gpurun_out/bench56.err; echo bench rc=$?
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,launch__grid_size
ncu --metrics $M --clock-control none -k regex:emd_solve_small -c 2 -s 8 --csv --log-file gpurun_out/k3_cont48.csv python bench.py --pairs 262144 --steps 1 --warmup 1 --no-cpu-baseline --no-table-arm > gpurun_out/ncu56.log 2>&1; echo rc=$?
