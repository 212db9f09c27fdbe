python -m pytest tests -m gpu -x -q > gpurun_out/pytest57.log 2>&1; echo pytest rc=$?; tail -2 gpurun_out/pytest57.log
python tools/stress_parity.py 100000 > gpurun_out/stress57.log 2>&1; echo stress rc=$?; tail -1 gpurun_out/stress57.log
python bench.py --mode sweep --steps 3 --lengths 32,64,128,256 > gpurun_out/sweep57.json 2> gpurun_out/sweep57.err; echo sweep rc=$?
