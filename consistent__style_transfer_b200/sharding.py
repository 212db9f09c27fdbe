"""Multi-GPU sharding of the WMD path: one process per GPU, pairs (or row blocks of an all-pairs
matrix) are independent, so ranks score disjoint slices with NO data-path collective and only the
final scores are gathered (BASELINE north_star "Only the final scores are
gathered, over NCCL/NVLink").  The reference itself is single-process, single-GPU
(/root/reference/job.yaml:29-32) -- this layer is new.

The embedding table and token maps are replicated per rank (12 MB at V=10k, d=300).  Slices are
contiguous in the caller's pair order and balanced by a per-pair cost model, so the gathered
output is in input order with no permutation step.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import numpy as np


def pair_cost(off1: np.ndarray, off2: np.ndarray) -> np.ndarray:
    """Relative cost of each pair: the cost tile is n1*n2 cells of d flops, the exact solve grows
    roughly with (n1+n2) pivots over n1*n2 cells, plus a fixed per-pair overhead."""
    n1 = np.diff(np.asarray(off1, np.int64)).astype(np.float64)
    n2 = np.diff(np.asarray(off2, np.int64)).astype(np.float64)
    return n1 * n2 + 8.0 * (n1 + n2) + 16.0


def partition(cost: np.ndarray, world: int) -> np.ndarray:
    """bounds[world + 1]: rank r owns items [bounds[r], bounds[r+1]); contiguous, monotone,
    every prefix as close as possible to r/world of the total cost."""
    n = int(cost.shape[0])
    bounds = np.zeros(world + 1, np.int64)
    bounds[world] = n
    if n == 0 or world == 1:
        return bounds
    csum = np.cumsum(cost)
    total = csum[-1]
    for r in range(1, world):
        bounds[r] = int(np.searchsorted(csum, total * r / world, side="left"))
    return np.maximum.accumulate(np.minimum(bounds, n))


def partition_by_tokens(off1: np.ndarray, off2: np.ndarray, world: int, per_pair: int = 16) -> np.ndarray:
    """``partition`` for the cost model "tokens of both sides + a per-pair constant" without touching the whole
    batch: the CSR offsets ARE the running token counts, so every bound is one binary search over
    f(p) = off1[p] + off2[p] + per_pair * p -- O(world * log B) instead of O(B) per call and rank."""
    n = int(off1.shape[0]) - 1
    bounds = np.zeros(world + 1, np.int64)
    bounds[world] = n
    if n == 0 or world == 1:
        return bounds
    base = int(off1[0]) + int(off2[0])
    f = lambda p: int(off1[p]) + int(off2[p]) - base + per_pair * p
    total = f(n)
    for r in range(1, world):
        want = total * r / world
        lo, hi = 0, n
        while lo < hi:                                   # smallest p with f(p) >= want
            mid = (lo + hi) // 2
            if f(mid) >= want:
                hi = mid
            else:
                lo = mid + 1
        bounds[r] = lo
    return np.maximum.accumulate(bounds)


def row_blocks(nrows: int, world: int) -> np.ndarray:
    """All-pairs mode: rank r owns rows [r*N/world, (r+1)*N/world) x all columns."""
    return (np.arange(world + 1, dtype=np.int64) * nrows) // world


def csr_slice(ids: np.ndarray, off: np.ndarray, lo: int, hi: int) -> Tuple[np.ndarray, np.ndarray]:
    off = np.asarray(off, np.int64)
    return np.ascontiguousarray(ids[off[lo]:off[hi]]), np.ascontiguousarray(off[lo:hi + 1] - off[lo])


def parse_cpulist(text: str) -> set:
    """'0-3,8,10-11' -> {0, 1, 2, 3, 8, 10, 11} (the format of sysfs local_cpulist / cpuset files)."""
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_host_to_gpu(device_index: int, sysfs: str = "/sys/bus/pci/devices") -> Optional[dict]:
    """One process per GPU: run this rank's host threads on the CPUs of the NUMA node its GPU hangs off, BEFORE pinned
    staging buffers are allocated (first touch puts them on that node).  Every rank of a multi-GPU host job copies its
    slice host -> device at the same time; with the pages on the far socket those copies cross the inter-socket link
    and the end-to-end step, unlike the kernels, stops scaling.  Returns what was done (None when the topology is not
    visible or the node's CPUs are outside this process's cpuset -- then nothing changes)."""
    import os
    try:
        import torch
        p = torch.cuda.get_device_properties(device_index)
        bdf = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(os.path.join(sysfs, bdf, "local_cpulist")) as f:
            local = parse_cpulist(f.read())
        with open(os.path.join(sysfs, bdf, "numa_node")) as f:
            node = int(f.read().strip())
        allowed = os.sched_getaffinity(0)
        cpus = local & allowed
        if not cpus or cpus == allowed:
            return {"bdf": bdf, "numa_node": node, "bound": False, "cpus": len(allowed)}
        os.sched_setaffinity(0, cpus)
        return {"bdf": bdf, "numa_node": node, "bound": True, "cpus": len(cpus)}
    except (OSError, ValueError, AttributeError, RuntimeError):
        return None


def _dist():
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist
    return None


def gather_scores(local_out, local_status, bounds: np.ndarray, group=None):
    """All-gather variable-length per-rank results (torch tensors on the backend's device: CUDA
    for NCCL, CPU for gloo) into full-length tensors present on every rank.  One collective of
    world * max_slice * 12 bytes -- latency-bound on NVLink."""
    import torch
    dist = _dist()
    world = len(bounds) - 1
    total = int(bounds[-1])
    if dist is None or world == 1:
        return local_out, local_status
    sizes = np.diff(bounds)
    m = int(sizes.max()) if total else 0
    dev = local_out.device
    pad_o = torch.zeros(m, dtype=torch.float64, device=dev)
    pad_s = torch.zeros(m, dtype=torch.int32, device=dev)
    pad_o[:local_out.numel()] = local_out
    pad_s[:local_status.numel()] = local_status
    all_o = torch.empty(world * m, dtype=torch.float64, device=dev)
    all_s = torch.empty(world * m, dtype=torch.int32, device=dev)
    if m:
        dist.all_gather_into_tensor(all_o, pad_o, group=group)
        dist.all_gather_into_tensor(all_s, pad_s, group=group)
    out = torch.empty(total, dtype=torch.float64, device=dev)
    st = torch.empty(total, dtype=torch.int32, device=dev)
    for r in range(world):
        lo, hi = int(bounds[r]), int(bounds[r + 1])
        out[lo:hi] = all_o[r * m:r * m + (hi - lo)]
        st[lo:hi] = all_s[r * m:r * m + (hi - lo)]
    return out, st


class PeerScores:
    """The global result of a sharded pair job, resident on every rank and filled by the ranks' KERNELS: each rank's
    pair entries store every score / status straight into the peers' copies over NVLink (``WMDEngine.set_fanout``),
    so the job needs no all-gather after the kernels -- the stores overlap the solves -- only a barrier before the
    result is read.  One process per GPU on one box (CUDA IPC); two buffer sets are used in turn, so a rank may start
    its next job while a peer still reads the previous result on its stream.

        peer = PeerScores(engine, total_pairs)                 # collective: every rank of the group
        out, st, (lo, hi) = wmd_pairs_sharded(engine.wmd_pairs_torch, ids1, off1, ids2, off2, peer=peer)
        peer.close()
    """

    SETS = 2

    def __init__(self, engine, total: int, group=None):
        import torch
        dist = _dist()
        if dist is None:
            raise RuntimeError("PeerScores needs an initialised torch.distributed process group")
        self.engine, self.total, self.group = engine, int(total), group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        if self.world - 1 > 7:
            raise RuntimeError("the fan-out holds at most 7 peers")
        self.dev = torch.device("cuda", engine.device)
        self.stride = ((12 * self.total + 255) // 256) * 256          # one set: float64 scores, then int32 status
        # Every step below is collective-safe: a rank that fails still takes part in the exchanges, and then all ranks raise.
        self.base, handle, err = 0, None, None
        try:
            self.base, handle = engine.peer_alloc(self.SETS * self.stride)
        except RuntimeError as exc:
            err = exc
        handles = [None] * self.world
        dist.all_gather_object(handles, handle, group=group)
        self.peer_base = [0] * self.world
        if err is None and all(h is not None for h in handles):
            try:
                for r in range(self.world):
                    self.peer_base[r] = self.base if r == self.rank else engine.peer_open(handles[r])
            except RuntimeError as exc:
                err = exc
        elif err is None:
            err = RuntimeError("a peer could not allocate its result buffer")
        oks = [None] * self.world
        dist.all_gather_object(oks, err is None, group=group)
        if not all(oks):
            for r, b in enumerate(self.peer_base):
                if b and r != self.rank:
                    engine.peer_close(b, True)
            dist.barrier(group=group)
            if self.base:
                engine.peer_close(self.base, False)
            self.peer_base = []
            raise RuntimeError(f"PeerScores: peer mapping unavailable ({err or 'on another rank'})")
        self._flag = torch.zeros(1, dtype=torch.int32, device=self.dev)
        self._turn = 0
        self._views = [self._wrap(k) for k in range(self.SETS)]
        dist.barrier(group=group)

    def _wrap(self, k: int):
        import torch

        class _Mem:                                                      # torch.as_tensor reads __cuda_array_interface__
            def __init__(self, ptr, shape, typestr):
                self.__cuda_array_interface__ = {"data": (ptr, False), "shape": shape, "typestr": typestr, "version": 2}
        b = self.base + k * self.stride
        out = torch.as_tensor(_Mem(b, (self.total,), "<f8"), device=self.dev)
        st = torch.as_tensor(_Mem(b + 8 * self.total, (self.total,), "<i4"), device=self.dev)
        return out, st

    def begin(self, lo: int):
        """Points the engine's fan-out at the peers' copies of the set in turn, advanced to this rank's first pair;
        returns this rank's own (out, status) tensors of that set."""
        k = self._turn
        outs = [b + k * self.stride + 8 * lo for r, b in enumerate(self.peer_base) if r != self.rank]
        sts = [b + k * self.stride + 8 * self.total + 4 * lo for r, b in enumerate(self.peer_base) if r != self.rank]
        self.engine.set_fanout(outs, sts)
        return self._views[k]

    def end(self):
        """Fan-out off, and the cross-rank barrier (stream-ordered: one tiny NCCL all-reduce) after which every rank's
        copy of the set is complete."""
        dist = _dist()
        self.engine.set_fanout()
        dist.all_reduce(self._flag, group=self.group)
        out, st = self._views[self._turn]
        self._turn = (self._turn + 1) % self.SETS
        return out, st

    def close(self):
        import torch
        dist = _dist()
        torch.cuda.synchronize(self.dev)
        if dist is not None:
            dist.barrier(group=self.group)
        self._views = []
        for r, b in enumerate(self.peer_base):
            if r != self.rank:
                self.engine.peer_close(b, True)
        if dist is not None:
            dist.barrier(group=self.group)                               # nobody still maps the buffer we free
        self.engine.peer_close(self.base, False)
        self.peer_base = []


def wmd_pairs_sharded(score_fn: Callable, ids1, off1, ids2, off2, group=None, gather: bool = True,
                      rank: Optional[int] = None, world: Optional[int] = None, balance: str = "tokens",
                      peer: Optional["PeerScores"] = None):
    """Every rank passes the SAME full CSR batch (host numpy); rank r scores its contiguous slice with
    ``score_fn(ids1, off1_slice, ids2, off2_slice) -> (float64 tensor, int32 tensor)`` (e.g.
    ``engine.wmd_pairs_torch``; the offset slices are views that keep pointing into the full id arrays, nothing is
    copied) and, with ``gather``, every rank receives all scores in input order.  ``balance``: "tokens" (slices of
    equal token count, found by binary search on the offsets) or "cost" (the quadratic per-pair model of
    ``pair_cost``, one pass over the batch).  With ``peer`` (a ``PeerScores`` of the batch's size) the gather is fused
    into the kernels: ``score_fn`` must accept ``out=`` / ``status=`` tensors (``engine.wmd_pairs_torch`` does) and the
    returned tensors are the rank's resident copy of the global result.  Returns (out, status, (lo, hi))."""
    dist = _dist()
    if world is None:
        world = dist.get_world_size(group) if dist else 1
    if rank is None:
        rank = dist.get_rank(group) if dist else 0
    off1 = np.asarray(off1, np.int64); off2 = np.asarray(off2, np.int64)
    bounds = partition_by_tokens(off1, off2, world) if balance == "tokens" else partition(pair_cost(off1, off2), world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    if peer is not None:
        g_out, g_st = peer.begin(lo)
        try:
            score_fn(ids1, off1[lo:hi + 1], ids2, off2[lo:hi + 1], out=g_out[lo:hi], status=g_st[lo:hi])
        finally:
            out, st = peer.end()
        return out, st, (lo, hi)
    out, st = score_fn(ids1, off1[lo:hi + 1], ids2, off2[lo:hi + 1])
    if gather:
        out, st = gather_scores(out, st, bounds, group=group)
    return out, st, (lo, hi)
