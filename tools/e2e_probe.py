"""Where does the host entry's extra time go?  Device-resident vs pinned-host calls at several batch sizes."""
import sys, time
import numpy as np, torch
sys.path.insert(0, ".")
from consistent__style_transfer_b200 import workload
from consistent__style_transfer_b200.engine import WMDEngine

table = workload.make_table(10000, 300, seed=0)
eng = WMDEngine(table)
dev = torch.device("cuda", 0)
for n in (65536, 262144, 1000000):
    ids1, off1, ids2, off2 = workload.make_pairs(n, "yelp", "independent", V=10000, seed=1)
    ml1, ml2 = int(np.diff(off1).max()), int(np.diff(off2).max())
    d = [torch.from_numpy(a).to(dev) for a in (ids1, off1, ids2, off2)]
    out = torch.empty(n, dtype=torch.float64, device=dev); st = torch.empty(n, dtype=torch.int32, device=dev)
    h = [torch.from_numpy(a).pin_memory() for a in (ids1, off1, ids2, off2)]
    ho = torch.empty(n, dtype=torch.float64).pin_memory(); hs = torch.empty(n, dtype=torch.int32).pin_memory()
    for _ in range(3):
        eng.wmd_pairs_cuda(d[0], d[1], d[2], d[3], ml1, ml2, out=out, status=st)
        eng.wmd_pairs_ptr(h[0].data_ptr(), h[1].data_ptr(), h[2].data_ptr(), h[3].data_ptr(), n, ho.data_ptr(), hs.data_ptr())
    torch.cuda.synchronize()
    td, th = [], []
    for _ in range(5):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        eng.wmd_pairs_cuda(d[0], d[1], d[2], d[3], ml1, ml2, out=out, status=st)
        torch.cuda.synchronize(); td.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        eng.wmd_pairs_ptr(h[0].data_ptr(), h[1].data_ptr(), h[2].data_ptr(), h[3].data_ptr(), n, ho.data_ptr(), hs.data_ptr())
        th.append(time.perf_counter() - t0)
    # raw copy cost of the same bytes
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for a, b in zip(h, d):
        b.copy_(a, non_blocking=True)
    ho.copy_(out, non_blocking=True); hs.copy_(st, non_blocking=True)
    torch.cuda.synchronize(); tc = time.perf_counter() - t0
    print(f"n={n}: device {1e3*np.median(td):.2f} ms, host {1e3*np.median(th):.2f} ms, plain copies {1e3*tc:.2f} ms")
