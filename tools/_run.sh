WMD_DTAB=1 python -m pytest tests -m gpu -x -q > gpurun_out/pytest50t.log 2>&1; echo pytest-dtab rc=$?; tail -2 gpurun_out/pytest50t.log
python -m pytest tests -m gpu -x -q > gpurun_out/pytest50.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/pytest50.log
WMD_DTAB=1 python tools/stress_parity.py 100000 > gpurun_out/stress50t.log 2>&1; echo stress-dtab rc=$?; tail -1 gpurun_out/stress50t.log
python bench.py --no-cpu-baseline > gpurun_out/bench50.json 2> gpurun_out/bench50.err; echo bench rc=$?
