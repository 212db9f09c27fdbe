python -m pytest tests -m gpu -x -q > gpurun_out/pytest48.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/pytest48.log
python tools/stress_parity.py 200000 > gpurun_out/stress48.log 2>&1; echo stress rc=$?; tail -1 gpurun_out/stress48.log
python bench.py --no-cpu-baseline > gpurun_out/bench48.json 2> gpurun_out/bench48.err; echo bench rc=$?
python bench.py --no-cpu-baseline --variant noised > gpurun_out/bench48n.json 2> gpurun_out/bench48n.err; echo bench rc=$?
