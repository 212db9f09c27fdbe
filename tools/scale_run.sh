#!/bin/bash
# 1/2/4/8-GPU scaling of the pair path and all-pairs mode, and the length sweep at 1 and 8 GPUs, on one box
# (run under `gpurun --gpus 8`); results in gpurun_out/scale_*.json
set -u
mkdir -p gpurun_out
port=29600
for n in 1 2 4 8; do
  port=$((port+1))
  if [ "$n" = 1 ]; then
    python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_pairs_$n.json 2> gpurun_out/scale_pairs_$n.err
    python bench.py --mode allpairs --docs 100000 --steps 3 --verify 0 > gpurun_out/scale_allpairs_$n.json 2> gpurun_out/scale_allpairs_$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 5 --warmup 3 2> gpurun_out/scale_pairs_$n.err | grep '"metric"' > gpurun_out/scale_pairs_$n.json
    port=$((port+1))
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --mode allpairs --docs 100000 --steps 3 --verify 0 2> gpurun_out/scale_allpairs_$n.err | grep '"metric"' > gpurun_out/scale_allpairs_$n.json
  fi
  if [ "$n" = 1 ] || [ "$n" = 8 ]; then
    port=$((port+1))
    if [ "$n" = 1 ]; then
      python bench.py --mode sweep --steps 3 > gpurun_out/scale_sweep_$n.json 2> gpurun_out/scale_sweep_$n.err
    else
      python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --mode sweep --steps 3 2> gpurun_out/scale_sweep_$n.err | grep '"metric"' > gpurun_out/scale_sweep_$n.json
    fi
  fi
  python - <<PY
import json
for m in ("pairs","allpairs"):
    try:
        j=json.load(open("gpurun_out/scale_%s_$n.json"%m)); print("$n", m, j["value"], j["ms_per_step"])
    except Exception as e: print("$n", m, "failed", e)
PY
done
