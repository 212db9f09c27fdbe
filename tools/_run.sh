python -m pytest tests -m gpu -x -q > gpurun_out/pytest42.log 2>&1; echo pytest rc=$?; tail -3 gpurun_out/pytest42.log
python tools/stress_parity.py 50000 > gpurun_out/stress42.log 2>&1; echo stress rc=$?; grep -c OK gpurun_out/stress42.log; grep MISMATCH gpurun_out/stress42.log
python bench.py > gpurun_out/bench42.json 2> gpurun_out/bench42.err; echo bench rc=$?
