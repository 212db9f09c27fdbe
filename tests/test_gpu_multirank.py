"""Two ranks on two GPUs over NCCL: the product's multi-GPU layer end to end (skips on a one-GPU box).

`sharding.wmd_pairs_sharded(engine.wmd_pairs_torch, ...)` -- token-balanced slices, every rank's library call on its own
device, NCCL gather of scores + status -- must give every rank the single-GPU result bit for bit; the all-pairs row
blocks computed by `allpairs_topk_cuda` and all-gathered on the device must equal the one-call result.
"""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    from consistent__style_transfer_b200 import sharding, workload
    from consistent__style_transfer_b200.engine import WMDEngine
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        V = 1200
        table = workload.make_table(V, 100, seed=3)
        eng = WMDEngine(table, device=rank)
        ids1, off1, ids2, off2 = workload.make_pairs(90_000, "uniform:1-40", "independent", V=V, seed=9)
        out, st, (lo, hi) = sharding.wmd_pairs_sharded(eng.wmd_pairs_torch, ids1, off1, ids2, off2)
        torch.cuda.synchronize()
        # the same job with the gather fused into the kernels: peer stores over NVLink, no all-gather; three jobs in a row
        # through the two buffer sets, host entry and device entry, documents above 32 tokens included (every kernel family)
        peer = sharding.PeerScores(eng, 90_000)
        p_out = p_st = None
        for it in range(3):
            if it == 1:                                                   # device-resident slice through wmd_pairs_cuda
                d = [torch.from_numpy(np.ascontiguousarray(x)).to(dev) for x in
                     (ids1[off1[lo]:off1[hi]], off1[lo:hi + 1] - off1[lo], ids2[off2[lo]:off2[hi]], off2[lo:hi + 1] - off2[lo])]
                fn = lambda a, b, c, e, out=None, status=None: eng.wmd_pairs_cuda(d[0], d[1], d[2], d[3], 40, 40, out=out, status=status)
            else:
                fn = eng.wmd_pairs_torch
            p_out, p_st, (plo, phi) = sharding.wmd_pairs_sharded(fn, ids1, off1, ids2, off2, peer=peer)
            assert (plo, phi) == (lo, hi)
            torch.cuda.synchronize()
            assert torch.equal(p_out, out) and torch.equal(p_st, st), f"peer gather differs in job {it}"
        peer.close()
        # all-pairs: this rank's row block, result left on the device, all-gathered over NCCL
        docs, doff, _, _ = workload.make_pairs(700, "yelp", "independent", V=V, seed=10)
        blocks = sharding.row_blocks(700, world)
        r0, r1 = int(blocks[rank]), int(blocks[rank + 1])
        mx = int(np.diff(blocks).max())
        k = 6
        t_idx = torch.zeros((mx, k), dtype=torch.int32, device=dev); t_dst = torch.zeros((mx, k), dtype=torch.float64, device=dev)
        eng.allpairs_topk_cuda(docs, doff, docs, doff, k, r0, r1, out_idx=t_idx[:r1 - r0], out_dist=t_dst[:r1 - r0])
        g_idx = torch.empty((world * mx, k), dtype=torch.int32, device=dev); g_dst = torch.empty((world * mx, k), dtype=torch.float64, device=dev)
        dist.all_gather_into_tensor(g_idx, t_idx); dist.all_gather_into_tensor(g_dst, t_dst)
        torch.cuda.synchronize()
        rows = [g_idx[r * mx:r * mx + int(blocks[r + 1] - blocks[r])].cpu().numpy() for r in range(world)]
        dsts = [g_dst[r * mx:r * mx + int(blocks[r + 1] - blocks[r])].cpu().numpy() for r in range(world)]
        q.put((rank, lo, hi, out.cpu().numpy().tobytes(), st.cpu().numpy().tobytes(),
               np.concatenate(rows).tobytes(), np.concatenate(dsts).tobytes()))
        eng.close()
    finally:
        dist.destroy_process_group()


def test_two_ranks_match_one_gpu():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    from consistent__style_transfer_b200 import workload
    from consistent__style_transfer_b200.engine import WMDEngine
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    V = 1200
    table = workload.make_table(V, 100, seed=3)
    eng = WMDEngine(table, device=0)
    ids1, off1, ids2, off2 = workload.make_pairs(90_000, "uniform:1-40", "independent", V=V, seed=9)
    want, wst = eng.wmd_pairs(ids1, off1, ids2, off2)
    docs, doff, _, _ = workload.make_pairs(700, "yelp", "independent", V=V, seed=10)
    widx, wdst, _ = eng.allpairs_topk(docs, doff, docs, doff, 6)
    eng.close()
    covered = 0
    for rank, lo, hi, ob, sb, ib, db in res:
        assert lo == covered and hi > lo
        covered = hi
        assert ob == want.tobytes() and sb == wst.tobytes()          # every rank holds all scores, input order
        assert ib == widx.tobytes() and db == wdst.tobytes()
    assert covered == 90_000
    # balanced in the partition's cost model (tokens of both sides + 16 per pair) to within one pair
    cost = lambda a, b: int(off1[b] - off1[a] + off2[b] - off2[a]) + 16 * (b - a)
    assert abs(cost(0, res[0][2]) - cost(res[0][2], 90_000)) <= 2 * (80 + 16)
