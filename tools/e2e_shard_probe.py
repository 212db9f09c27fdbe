"""Where the sharded end-to-end step spends its time (run under torchrun with 2+ ranks)."""
import os, sys, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from consistent__style_transfer_b200 import sharding, workload
from consistent__style_transfer_b200.engine import WMDEngine

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local); dev = torch.device("cuda", local)
numa = sharding.bind_host_to_gpu(local) if os.environ.get("WMD_NUMA_BIND", "1") != "0" else None
dist.init_process_group("nccl", device_id=dev)
table = workload.make_table(10000, 300, seed=0)
eng = WMDEngine(table, device=local)
ids1, off1, ids2, off2 = workload.make_pairs(world * 1_000_000, "yelp", "independent", V=10000, seed=1)
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
ids1, off1, ids2, off2 = pin(ids1), pin(off1), pin(ids2), pin(off2)
n = world * 1_000_000
h_out = torch.empty(n, dtype=torch.float64).pin_memory(); h_st = torch.empty(n, dtype=torch.int32).pin_memory()
T = {}
def tick(name, t0):
    torch.cuda.synchronize(); T[name] = T.get(name, 0.0) + time.perf_counter() - t0
for it in range(8):
    if it == 3: T.clear()
    dist.barrier(); torch.cuda.synchronize()
    t_all = time.perf_counter()
    t0 = time.perf_counter(); bounds = sharding.partition_by_tokens(off1, off2, world); lo, hi = int(bounds[rank]), int(bounds[rank + 1]); tick("partition", t0)
    t0 = time.perf_counter(); out, st = eng.wmd_pairs_torch(ids1, off1[lo:hi + 1], ids2, off2[lo:hi + 1]); tick("score", t0)
    t0 = time.perf_counter(); g_out, g_st = sharding.gather_scores(out, st, bounds); tick("gather", t0)
    t0 = time.perf_counter(); h_out[lo:hi].copy_(g_out[lo:hi], non_blocking=True); h_st[lo:hi].copy_(g_st[lo:hi], non_blocking=True); tick("d2h", t0)
    T["total"] = T.get("total", 0.0) + time.perf_counter() - t_all
res = [None] * world
dist.all_gather_object(res, ({k: round(v / 5 * 1e3, 3) for k, v in T.items()}, numa))
if rank == 0:
    for r, x in enumerate(res):
        print(r, x)
dist.destroy_process_group()
