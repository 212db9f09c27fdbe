"""world_size-2 (and 3) gloo runs of the multi-rank host logic on CPU: cost-balanced contiguous
partition, per-rank scoring of disjoint slices, one gather of the scores in input order.  There is
no GPU here, so the per-rank scorer is the oracle -- what is under test is the sharding layer."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from consistent__style_transfer_b200 import sharding, workload


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, B, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import wmd_oracle
        table = workload.make_table(300, 16, seed=1)
        ids1, off1, ids2, off2 = workload.make_pairs(B, "book", "independent", V=300, seed=4)
        calls = []

        def score(a1, o1, a2, o2):
            calls.append(len(o1) - 1)
            out, st = wmd_oracle.batch_wmd(table, a1, o1, a2, o2)
            return torch.from_numpy(out), torch.from_numpy(st)

        out, st, (lo, hi) = sharding.wmd_pairs_sharded(score, ids1, off1, ids2, off2)

        # the peer= path of the same call (on GPUs: sharding.PeerScores, the kernels store into the peers' copies).  The
        # stand-in keeps the protocol -- begin(lo) hands out this rank's full-size result, the scorer fills its slice
        # in place, end() is the barrier -- and plays the peers' stores with one all-reduce of the disjoint slices.
        class FakePeer:
            def __init__(self, total):
                self.out = torch.zeros(total, dtype=torch.float64); self.st = torch.zeros(total, dtype=torch.int32)
                self.began = None

            def begin(self, lo_):
                self.began = lo_
                self.out.zero_(); self.st.zero_()
                return self.out, self.st

            def end(self):
                fin = torch.isfinite(self.out)
                inf_mask = (~fin).to(torch.int32)                          # +inf scores travel as a mask: the sum stays exact
                dist.all_reduce(inf_mask); dist.all_reduce(self.st)
                o = torch.where(fin, self.out, torch.zeros_like(self.out)); dist.all_reduce(o)
                self.out = torch.where(inf_mask > 0, torch.full_like(o, float("inf")), o)
                return self.out, self.st

        def score_into(a1, o1, a2, o2, out=None, status=None):
            o, s_ = score(a1, o1, a2, o2)
            out.copy_(o); status.copy_(s_)
            return out, status

        peer = FakePeer(B)
        p_out, p_st, (plo, phi) = sharding.wmd_pairs_sharded(score_into, ids1, off1, ids2, off2, peer=peer)
        assert (plo, phi) == (lo, hi) and peer.began == lo
        assert torch.equal(p_out, out) and torch.equal(p_st, st)
        q.put((rank, lo, hi, calls[0], out.numpy().tobytes(), st.numpy().tobytes()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,B", [(2, 257), (3, 100), (2, 1)])
def test_sharded_scores_equal_single_rank(oracle, world, B):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, B, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    table = workload.make_table(300, 16, seed=1)
    ids1, off1, ids2, off2 = workload.make_pairs(B, "book", "independent", V=300, seed=4)
    want, wst = oracle.batch_wmd(table, ids1, off1, ids2, off2)
    covered = 0
    for rank, lo, hi, ncalls, ob, sb in res:
        assert ncalls == hi - lo                       # each rank scored only its own slice
        assert lo == covered
        covered = hi
        assert ob == want.tobytes() and sb == wst.tobytes()   # every rank holds all scores, input order
    assert covered == B


def test_partition_is_balanced_and_total():
    rng = np.random.default_rng(0)
    cost = rng.integers(1, 500, size=10_000).astype(np.float64)
    for world in (1, 2, 4, 8):
        b = sharding.partition(cost, world)
        assert b[0] == 0 and b[-1] == len(cost) and np.all(np.diff(b) >= 0)
        per = [cost[b[r]:b[r + 1]].sum() for r in range(world)]
        assert max(per) <= cost.sum() / world + cost.max()
    assert list(sharding.partition(np.zeros(0), 4)) == [0, 0, 0, 0, 0]
    assert list(sharding.row_blocks(10, 4)) == [0, 2, 5, 7, 10]


def test_csr_slice_rebases_offsets():
    ids, off = workload.to_csr([[1, 2], [], [3], [4, 5, 6]])
    a, o = sharding.csr_slice(ids, off, 1, 4)
    assert list(a) == [3, 4, 5, 6] and list(o) == [0, 0, 1, 4]


def test_partition_by_tokens_matches_the_linear_cost_model():
    ids1, off1, ids2, off2 = workload.make_pairs(20_000, "uniform:1-90", "independent", V=500, seed=3)
    cost = (np.diff(off1) + np.diff(off2) + 16).astype(np.float64)
    for world in (1, 2, 3, 8):
        b = sharding.partition_by_tokens(off1, off2, world)
        assert b[0] == 0 and b[-1] == 20_000 and np.all(np.diff(b) >= 0)
        per = [cost[b[r]:b[r + 1]].sum() for r in range(world)]
        assert max(per) - min(per) <= 2 * cost.max() + 1
        # a view that starts in the middle of the batch partitions its own pairs
        v = sharding.partition_by_tokens(off1[5000:15001], off2[5000:15001], world)
        assert v[-1] == 10_000 and np.all(np.diff(v) >= 0)
    assert list(sharding.partition_by_tokens(np.zeros(1, np.int64), np.zeros(1, np.int64), 3)) == [0, 0, 0, 0]


def test_parse_cpulist_and_binding_without_a_gpu():
    assert sharding.parse_cpulist("0-3,8,10-11\n") == {0, 1, 2, 3, 8, 10, 11}
    assert sharding.parse_cpulist("") == set()
    # no CUDA device / no sysfs entry: nothing changes, nothing raises
    import os
    before = os.sched_getaffinity(0)
    assert sharding.bind_host_to_gpu(0, sysfs="/nonexistent") is None
    assert os.sched_getaffinity(0) == before
