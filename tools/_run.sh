python -m pytest tests -m gpu -x -q > gpurun_out/pytest46.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/pytest46.log
python tools/stress_parity.py 100000 > gpurun_out/stress46.log 2>&1; echo stress rc=$?; tail -1 gpurun_out/stress46.log
python bench.py --no-cpu-baseline > gpurun_out/bench46.json 2> gpurun_out/bench46.err; echo bench rc=$?
