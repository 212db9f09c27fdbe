#!/usr/bin/env python
"""bench.py -- WMD sentence-pairs/s on the Yelp-shape synthetic batch (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's B200 engine
    python bench.py --impl reference [--gpus N] ...                # the reference's CPU path (oracle port)

A "step" is one pass of the hot path (nBOW -> word distances -> exact EMD) over one batch of synthetic pairs:
1 M Yelp-shape pairs per GPU (len 1..20, d=300, V=10k, `independent` variant = worst case).
Weak scaling: the global batch holds N x 1 M pairs, `sharding.partition_by_tokens` deals every rank a contiguous
slice of equal token count, and nothing but the final scores crosses NCCL (`sharding.gather_scores`, inside the timed
region; pairs are independent -- BASELINE north_star / SURVEY 8(e)).

value  = pairs/s with the rank's slice already resident in HBM (wmd_pairs_dev), CUDA-event timed, max over ranks.
e2e    = pairs/s through the public call on HOST buffers: N = 1 `WMDEngine.wmd_pairs` (wmd_pairs_host) on pinned
         arrays, N > 1 `sharding.wmd_pairs_sharded(engine.wmd_pairs_torch, ...)` on the global batch (every rank ends
         with all scores on its device) + a device->host read of the rank's own slice; H2D of ids + offsets and D2H of
         scores + status inside the timed region.
The engine runs its default policy: the V x V word-distance table (400 MB at V = 10k) is built ONCE by the first call
(warm-up; `table_build_ms`, `table_break_even_pairs`) and every step takes its costs from it through the fused
warp-per-pair kernel; `without_table` is the same measurement on the direct path that recomputes every float32
distance from the embedding rows, as gensim does.  Both produce identical bits.
roofline = the dominant kernel against the HBM roofline with SURVEY 8(d)'s algorithmic bytes (the contract's
         convention), next to what actually binds it: instruction issue (`issue`, `binding`), from the ncu capture of
         the same kernel kept under profiles/ (`traffic_source` names file and commit).
cpu_baseline / --impl reference = oracle/wmd_oracle.py's gensim-shaped python loop + C emd_hat (the reference itself
         cannot be installed: gensim and pyemd are absent).
Sub-records of the default run (one JSON line): `allpairs` (configs[3], strong scaling), `sweep` (configs[4] plus a
mixed-length workload), `latency` (configs[2]: one collate batch through the reference's own call surface; N = 1).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from consistent__style_transfer_b200 import workload  # noqa: E402

METRIC = "wmd_sentence_pairs_per_sec"
UNIT = "pairs/s"
NCU_FILE = os.path.join(ROOT, "profiles", "r02_ncu_fused.json")     # written by tools/ncu_summary.py from the .ncu-rep


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=1_000_000, help="pairs per GPU per step")
    ap.add_argument("--shape", default="yelp")
    ap.add_argument("--variant", default="independent", choices=["independent", "noised"])
    ap.add_argument("--d", type=int, default=300)
    ap.add_argument("--vocab", type=int, default=10_000)
    ap.add_argument("--cpu-sample", type=int, default=0, help="pairs in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-direct-arm", action="store_true", help="skip the additional measurement without the word-distance table")
    ap.add_argument("--no-extras", action="store_true", help="headline only: no allpairs / sweep / latency sub-records")
    ap.add_argument("--mode", default="pairs", choices=["pairs", "allpairs", "latency", "sweep"],
                    help="pairs = BASELINE configs[1] (the driver's headline, with the other modes as sub-records); "
                         "allpairs / sweep / latency = that mode alone")
    ap.add_argument("--lengths", default="8,16,32,64,128,256", help="sweep: document lengths")
    ap.add_argument("--docs", type=int, default=100_000, help="allpairs: documents in the set (self join)")
    ap.add_argument("--topk", type=int, default=16)
    ap.add_argument("--verify", type=int, default=2000, help="allpairs: brute-force check of the first N x N block (0 = off)")
    return ap.parse_args()


def workload_name(a):
    return f"{a.shape}-shape {a.variant} pairs, len<=20, d={a.d}, V={a.vocab}, {a.pairs} pairs/GPU/step"


def config_dict(a, n_gpus):
    """The same dictionary in both arms (the reference arm times a bounded sample of this workload per step)."""
    return {"workload": workload_name(a), "pairs_per_gpu": a.pairs, "global_pairs": n_gpus * a.pairs,
            "cost": "float32 numpy-order (bit-exact)", "emd": "pyemd 1e6-grid integer optimum, exact",
            "l2": "256 MiB flush write between timed steps",
            "parallelism": f"pairs sharded over {n_gpus} GPU(s) in token-balanced contiguous slices; no data-path "
                           f"collective; every rank ends each step with all float64 scores + status on its device, inside the "
                           f"timed region (stored by the kernels into the peers' copies over NVLink + a barrier, or one NCCL all-gather)"}


# ------------------------------------------------------------------------------------------------
# CPU side: the oracle port of the reference path, run in worker processes (one per core)
# ------------------------------------------------------------------------------------------------
_W = {}


def _cpu_init(table, ids1, off1, ids2, off2):
    from oracle import wmd_oracle
    words = ["w%06d" % i for i in range(table.shape[0])]      # zero-padded: string order == row order
    _W.update(kv=wmd_oracle.KeyedVectorsOracle(words, table), words=words, ids1=ids1, off1=off1, ids2=ids2, off2=off2)
    wmd_oracle.lib()


def _cpu_worker(span):
    lo, hi = span
    kv, words, ids1, off1, ids2, off2 = (_W[k] for k in ("kv", "words", "ids1", "off1", "ids2", "off2"))
    t0 = time.perf_counter()
    acc = 0.0
    for p in range(lo, hi):
        d1 = [words[t] for t in ids1[off1[p]:off1[p + 1]]]
        d2 = [words[t] for t in ids2[off2[p]:off2[p + 1]]]
        v = kv.wmdistance(d1, d2)                                 # the reference's per-pair call (src/wmd.py:32)
        if v != float("inf"):
            acc += v
    return time.perf_counter() - t0, hi - lo, acc


class CpuPool:
    """Worker processes (fork, one per core) running the reference-shaped python loop."""

    def __init__(self, table, pairs, cores):
        import multiprocessing as mp
        self.cores = cores
        self.pool = mp.get_context("fork").Pool(cores, initializer=_cpu_init, initargs=(table,) + tuple(pairs))
        self.pool.map(_cpu_worker, [(0, 1)] * cores)             # touch every worker once

    def run(self, lo, hi):
        bounds = np.linspace(lo, hi, self.cores + 1).astype(int)
        t0 = time.perf_counter()
        self.pool.map(_cpu_worker, [(int(bounds[i]), int(bounds[i + 1])) for i in range(self.cores)], chunksize=1)
        return time.perf_counter() - t0

    def close(self):
        self.pool.close()
        self.pool.join()


def cpu_c_port(table, pairs, n_sample, cores):
    from oracle import wmd_oracle
    ids1, off1, ids2, off2 = pairs
    n = min(n_sample, len(off1) - 1)
    a1, o1 = ids1[:off1[n]], off1[:n + 1]
    a2, o2 = ids2[:off2[n]], off2[:n + 1]
    wmd_oracle.batch_wmd(table, a1[:o1[64]], o1[:65], a2[:o2[64]], o2[:65], nthreads=cores)
    t0 = time.perf_counter()
    wmd_oracle.batch_wmd(table, a1, o1, a2, o2, nthreads=cores)
    return n / (time.perf_counter() - t0)


def cpu_one_core(table, pairs, n_sample):
    """The reference's real configuration: one process, one thread (src/main_pretrain.py:120-122, num_workers unset)."""
    _cpu_init(table, *pairs)
    _cpu_worker((0, 8))
    dt, n, _ = _cpu_worker((0, n_sample))
    return n / dt


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    n_gpus = max(a.gpus, world)
    cores = os.cpu_count() or 1
    table = workload.make_table(a.vocab, a.d, seed=0)
    per_step = a.cpu_sample or 1000 * cores
    pairs = workload.make_pairs(per_step * (a.steps + a.warmup), a.shape, a.variant, V=a.vocab, seed=1)
    pool = CpuPool(table, pairs, cores)
    times = []
    for s_ in range(a.warmup + a.steps):
        dt = pool.run(s_ * per_step, (s_ + 1) * per_step)
        if s_ >= a.warmup:
            times.append(dt)
    pool.close()
    total = sum(times)
    value = per_step * a.steps / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": a.steps,
        "warmup": a.warmup, "ms_per_step": 1e3 * total / a.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(a, n_gpus),
        "arm": {"sample_pairs_per_step": per_step,
                "note": "the reference cannot be installed (gensim / pyemd absent): this is the oracle port of its CPU path, "
                        "timed on a bounded sample of the configured workload per step and reported as a rate; "
                        "liboracle_wmd.so (the C emd_hat) is loaded inside the forked pool workers, so a hook that lists "
                        "the shared objects of THIS process does not see it"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{per_step} pairs/step of the same workload, python per-pair loop + C emd_hat, "
                                   f"{cores} worker processes, wall time per step"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU side
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.rows = []
        self._halt = threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [x.strip() for x in out.strip().split(",")]
                if len(parts) >= 7:
                    self.rows.append(parts)
            except Exception:
                pass
            self._halt.wait(0.1)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        sm = [float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({n for r in self.rows for n, v in zip(names, r[3:7]) if v.lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.rows)}


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class Ctx:
    """Process-wide state shared by the arms: ranks, device, process group, table, engine."""

    def __init__(self, a):
        import torch
        import torch.distributed as dist
        self.a = a
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.n_gpus = max(a.gpus, self.world)
        self.table = None
        self.eng = None
        self.dev = None
        self.flush = None

    def start_gpu(self, d=None):
        from consistent__style_transfer_b200.engine import WMDEngine
        torch = self.torch
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        self.numa = None
        if self.world > 1 and os.environ.get("WMD_NUMA_BIND", "1") != "0":
            from consistent__style_transfer_b200 import sharding
            self.numa = sharding.bind_host_to_gpu(self.local)          # before any pinned buffer exists
        if self.world > 1 and not self.dist.is_initialized():
            self.dist.init_process_group("nccl", device_id=self.dev)
        self.table = workload.make_table(self.a.vocab, d or self.a.d, seed=0)
        self.eng = WMDEngine(self.table, device=self.local)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)          # > 126 MB L2

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, x):
        t = self.torch.tensor([float(x)], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def close(self):
        if getattr(self, "peer", None) is not None:
            self.peer.close()
            self.peer = None
        if self.eng is not None:
            self.eng.close()
        if self.world > 1 and self.dist.is_initialized():
            self.dist.destroy_process_group()


def timed_device_steps(ctx, step, steps, warmup):
    """`warmup` untimed + `steps` timed calls of step(), L2 flushed before each; returns summed ms (this rank)."""
    torch = ctx.torch
    for _ in range(warmup):
        ctx.flush.fill_(1)
        step()
    ctx.barrier()
    evs = []
    for _ in range(steps):
        ctx.flush.fill_(1)                                                  # L2 flush, outside the event pair
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        evs.append((e0, e1))
    ctx.barrier()
    return sum(e0.elapsed_time(e1) for e0, e1 in evs)


def pairs_arm(ctx, cpu):
    from consistent__style_transfer_b200 import sharding
    a, torch, eng, dev, world, rank = ctx.a, ctx.torch, ctx.eng, ctx.dev, ctx.world, ctx.rank
    n_gpus = ctx.n_gpus
    # the global batch (the same on every rank) and this rank's token-balanced contiguous slice of it
    ids1, off1, ids2, off2 = workload.make_pairs(world * a.pairs, a.shape, a.variant, V=a.vocab, seed=1)
    bounds = sharding.partition_by_tokens(off1, off2, world)
    lo, hi = int(bounds[rank]), int(bounds[rank + 1])
    nloc = hi - lo
    l_ids1, l_off1 = sharding.csr_slice(ids1, off1, lo, hi)
    l_ids2, l_off2 = sharding.csr_slice(ids2, off2, lo, hi)
    ml1, ml2 = int(np.diff(l_off1).max()), int(np.diff(l_off2).max())
    d_ids1 = torch.from_numpy(l_ids1).to(dev); d_off1 = torch.from_numpy(l_off1).to(dev)
    d_ids2 = torch.from_numpy(l_ids2).to(dev); d_off2 = torch.from_numpy(l_off2).to(dev)
    d_out = torch.empty(nloc, dtype=torch.float64, device=dev)
    d_st = torch.empty(nloc, dtype=torch.int32, device=dev)
    gathered = {}

    # N > 1: the gather of the final scores (12 B/pair, the path's only exchange) is fused into the kernels -- every
    # rank's kernels store their scores into the peers' copies of the global result over NVLink (sharding.PeerScores),
    # a tiny all-reduce is the barrier.  WMD_PEER_GATHER=0 (or no CUDA IPC on the box): one NCCL all-gather instead.
    peer = None
    if world > 1 and os.environ.get("WMD_PEER_GATHER", "1") != "0":
        try:
            peer = sharding.PeerScores(eng, int(bounds[-1]))
        except Exception as exc:                                           # noqa: BLE001 -- any failure: the NCCL path
            sys.stderr.write(f"[bench] peer gather unavailable ({exc}); using the NCCL all-gather\n")
    if world > 1:                                                          # every rank must take the same path
        ok = torch.tensor([1 if peer is not None else 0], device=dev)
        ctx.dist.all_reduce(ok, op=ctx.dist.ReduceOp.MIN)
        if int(ok.item()) == 0 and peer is not None:
            peer.close(); peer = None
    ctx.peer = peer

    def dev_step():
        if peer is not None:
            g_out, g_st = peer.begin(lo)
            eng.wmd_pairs_cuda(d_ids1, d_off1, d_ids2, d_off2, ml1, ml2, out=g_out[lo:hi], status=g_st[lo:hi])
            gathered["out"], gathered["st"] = peer.end()
            return
        eng.wmd_pairs_cuda(d_ids1, d_off1, d_ids2, d_off2, ml1, ml2, out=d_out, status=d_st)
        if world > 1:                                                      # the only exchange: final scores + status, 12 B/pair
            gathered["out"], gathered["st"] = sharding.gather_scores(d_out, d_st, bounds)

    # pinned host copies for the end-to-end arm (N = 1: the whole batch is this rank's slice)
    h_ids1 = torch.from_numpy(ids1).pin_memory(); h_off1 = torch.from_numpy(off1).pin_memory()
    h_ids2 = torch.from_numpy(ids2).pin_memory(); h_off2 = torch.from_numpy(off2).pin_memory()
    n_all = world * a.pairs
    h_out = torch.empty(n_all, dtype=torch.float64).pin_memory()
    h_st = torch.empty(n_all, dtype=torch.int32).pin_memory()
    np_ids1, np_off1, np_ids2, np_off2 = h_ids1.numpy(), h_off1.numpy(), h_ids2.numpy(), h_off2.numpy()
    np_out, np_st = h_out.numpy(), h_st.numpy()

    def host_step():
        if world == 1:
            eng.wmd_pairs(np_ids1, np_off1, np_ids2, np_off2, out=np_out, status=np_st)       # returns after the D2H of the scores
        else:
            # every rank ends with ALL scores on its device (the product's contract) and reads its OWN slice back:
            # the global result reaches host memory exactly once per step
            out, st, (a0, a1) = sharding.wmd_pairs_sharded(eng.wmd_pairs_torch, np_ids1, np_off1, np_ids2, np_off2, peer=peer)
            h_out[a0:a1].copy_(out[a0:a1], non_blocking=True); h_st[a0:a1].copy_(st[a0:a1], non_blocking=True)
            gathered["e2e_out"] = out
            torch.cuda.synchronize()

    def measure(label):
        t0 = time.perf_counter()
        dev_step()                                                         # the first call of an engine builds its table
        torch.cuda.synchronize()
        first_s = time.perf_counter() - t0
        eng.set_profiling(False)
        for _ in range(a.warmup):
            ctx.flush.fill_(1)
            dev_step()
        ctx.barrier()
        eng.set_profiling(True); eng.profile(reset=True)
        total_ms = timed_device_steps(ctx, dev_step, a.steps, 0)
        prof = eng.profile(reset=True)
        eng.set_profiling(False)
        stats = eng.last_stats()
        # one more profiled pass with the chunks serialised on a single stream: the kernels' own durations
        eng.set_serial(True); eng.set_profiling(True); eng.profile(reset=True)
        for _ in range(2):
            ctx.flush.fill_(1)
            dev_step()
        ctx.barrier()
        prof_serial = eng.profile(reset=True)
        eng.set_profiling(False); eng.set_serial(False)
        total_ms_max = ctx.max_over_ranks(total_ms)
        dev_snapshot = (gathered["out"] if world > 1 else d_out).clone()   # the peer buffer sets are reused by the e2e arm
        for _ in range(a.warmup):
            host_step()
        ctx.barrier()
        e2e_s = 0.0
        for _ in range(a.steps):
            ctx.flush.fill_(1)
            ctx.barrier()
            t0 = time.perf_counter()
            host_step()
            e2e_s += time.perf_counter() - t0
        ctx.barrier()
        e2e_s = ctx.max_over_ranks(e2e_s)
        dev_scores = dev_snapshot.cpu().numpy()
        if world > 1:                                                      # the e2e arm's gathered tensor against the device arm's
            same = bool(np.array_equal(gathered["e2e_out"].cpu().numpy(), dev_scores)) and bool(np.array_equal(np_out[lo:hi], dev_scores[lo:hi]))
        else:
            same = bool(np.array_equal(np_out, dev_scores))
        return {"label": label, "total_ms": total_ms, "total_ms_max": total_ms_max, "prof": prof, "prof_serial": prof_serial,
                "stats": stats, "e2e_s": e2e_s, "first_call_s": first_s,
                "value": n_all * a.steps / (total_ms_max / 1e3), "e2e_value": n_all * a.steps / e2e_s,
                "same": same, "scores": dev_scores}

    sampler = ClockSampler(ctx.local)
    sampler.start()
    main = measure("default")
    clocks = sampler.stop()
    tinfo = eng.distance_table_info()
    direct = None
    if not a.no_direct_arm:
        eng.set_distance_table(False)
        direct = measure("direct")
        eng.set_distance_table(tinfo["enabled"])

    # ---- roofline of the dominant kernel --------------------------------------------------------
    peak, peak_src = hbm_peak()
    stats = main["stats"]
    alg_bytes_step = 4 * stats["tokens"] + 4 * a.d * stats["uniques"] + 8 * nloc
    table_bytes_step = 4 * stats["tokens"] + 4 * stats["cells"] + 8 * nloc
    kern = {k: v for k, v in main["prof"].items() if v["launches"] > 0}
    dom = max(kern, key=lambda k: kern[k]["ms"])
    dom_ms_step = kern[dom]["ms"] / a.steps
    serial = {k: v["ms"] / 2 for k, v in main["prof_serial"].items() if v["launches"] > 0}
    nlaunch = max(1, kern[dom]["launches"] // a.steps)
    ncu = json.load(open(NCU_FILE)) if os.path.exists(NCU_FILE) else None
    sm_clock = (clocks.get("sm_mhz") or 1965.0) * 1e6
    roofline = {
        "bound": "hbm", "kernel": dom, "achieved": alg_bytes_step / (dom_ms_step / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
        "frac": alg_bytes_step / (dom_ms_step / 1e3) / 1e9 / peak,
        "traffic": (ncu or {}).get("dram_bytes_per_launch"),
        "traffic_source": (f"{os.path.relpath(NCU_FILE, ROOT)}: {ncu.get('source')} at commit {ncu.get('commit')}, "
                           f"{ncu.get('pairs_per_launch')} pairs per launch" if ncu else None),
        "peak_source": peak_src,
        "algorithmic_bytes_per_step": alg_bytes_step, "algorithmic_bytes_per_launch": alg_bytes_step / nlaunch,
        "binding": "instruction issue: the table mode moves 4 B per cost cell instead of d floats per row, so neither HBM nor "
                   "the FP32 pipe limits the fused kernel -- the exact transport solve does (data-dependent pivots, REDUX / "
                   "ballot / shuffle per selection); `issue` is the kernel's share of issue slots in use",
        "issue": ({"issue_active_pct": ncu.get("issue_active_pct"), "warp_instructions_per_pair": ncu.get("warp_instructions_per_pair"),
                   "ceiling_pairs_per_s_at_100pct_issue": (148 * 4 * sm_clock / ncu["warp_instructions_per_pair"]
                                                           if ncu.get("warp_instructions_per_pair") else None)} if ncu else None),
        "table_mode": {"bytes_per_step": table_bytes_step, "note": "what the fused kernel really has to fetch: ids + 4 B per cell of "
                       "the u1 x u2 tile from the word-distance table + the score", "achieved_gbs": table_bytes_step / (dom_ms_step / 1e3) / 1e9},
        "note": "achieved = SURVEY 8(d) algorithmic bytes of the step (4(n1+n2) + 4d(u1+u2) + 8 per pair, from the engine's own counters) / "
                "summed CUDA-event time of the dominant kernel's launches in the timed steps, recorded on the kernels' own streams "
                "(chunks on two streams overlap, so these times include co-scheduling; `standalone` has the serialised durations)",
        "kernel_ms_per_step": {k: v["ms"] / a.steps for k, v in kern.items()},
        "launches_per_step": {k: v["launches"] // a.steps for k, v in kern.items()},
        "standalone": {"note": "same step with all chunks on one stream (2 extra steps outside the headline timing)",
                       "kernel_ms_per_step": serial,
                       "achieved": alg_bytes_step / (serial[dom] / 1e3) / 1e9, "frac": alg_bytes_step / (serial[dom] / 1e3) / 1e9 / peak},
        "whole_step_achieved": alg_bytes_step / (main["total_ms"] / a.steps / 1e3) / 1e9}
    if direct is not None:
        dk = {k: v["ms"] / 2 for k, v in direct["prof_serial"].items() if v["launches"] > 0}
        cells = stats["cells"] / max(1, world)                                # this rank's cells per step
        cost_ms = dk.get("cost")
        roofline["direct_path"] = {
            "kernel_ms_per_step_standalone": dk,
            "cost_hbm_frac": alg_bytes_step / (cost_ms / 1e3) / 1e9 / peak if cost_ms else None,
            "cost_fp32_frac": (direct["stats"]["cells"] * 3 * a.d / (cost_ms / 1e3) / (148 * 128 * sm_clock) if cost_ms else None),
            "note": "direct path (no table): the cost kernel's FP32 pipe floor is cells x 3d separately rounded operations "
                    "(numpy's sub, square, add) / (SMs x 128 lanes x clock)"}
    launches = sum(v["launches"] for v in kern.values())
    line = None
    if rank == 0:
        h2d = int(ids1.nbytes + ids2.nbytes + off1.nbytes + off2.nbytes) // (1 if world == 1 else 1)
        line = {
            "metric": METRIC, "value": main["value"], "unit": UNIT, "n_gpus": n_gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": main["total_ms_max"] / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "data": "synthetic",
            "config": config_dict(a, n_gpus),
            "clocks": clocks,
            "e2e": {"value": main["e2e_value"], "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": int(n_all * 12),
                    "timing": "perf_counter around the public call on pinned host buffers (N = 1: WMDEngine.wmd_pairs; N > 1: "
                              "sharding.wmd_pairs_sharded -- H2D of the rank's slice, kernels, NCCL gather of all scores on every "
                              "device -- + device->host read of the rank's own slice, so the global result reaches host memory once), "
                              "max over ranks",
                    "matches_device_path": main["same"],
                    "host_binding": ctx.numa},                            # N > 1: rank 0's CPU affinity / NUMA node (sharding.bind_host_to_gpu)
            "gpu_launches": launches,
            "score_gather": ("none (one GPU)" if world == 1 else
                             "fused: the kernels store scores + status into the peers' copies over NVLink (CUDA IPC), barrier = one 4-byte all-reduce"
                             if ctx.peer is not None else "one NCCL all-gather of scores + status after the kernels"),
            "roofline": roofline,
            "word_distance_table": {"enabled": tinfo["enabled"], "bytes": tinfo["bytes"], "table_build_ms": tinfo["build_ms"],
                                    "first_call_s": main["first_call_s"],
                                    "note": "default policy: built once per embedding table by the first scoring call (here: before "
                                            "the warm-up steps), never inside a timed step"},
            "mean_len": float(stats["tokens"]) / (2 * max(nloc, 1)),
        }
        if direct is not None:
            per_pair_gain_ms = (direct["total_ms_max"] - main["total_ms_max"]) / a.steps / a.pairs
            line["without_table"] = {
                "value": direct["value"], "unit": UNIT, "ms_per_step": direct["total_ms_max"] / a.steps, "e2e_value": direct["e2e_value"],
                "bit_identical_to_default": bool(np.array_equal(direct["scores"], main["scores"])),
                "kernel_ms_per_step": {k: v["ms"] / a.steps for k, v in direct["prof"].items() if v["launches"] > 0},
                "note": "wmd_set_distance_table(0): every float32 distance recomputed from the embedding rows for every pair, as gensim does"}
            line["word_distance_table"]["table_break_even_pairs"] = (tinfo["build_ms"] / per_pair_gain_ms if per_pair_gain_ms > 0 else None)
        if cpu is not None:
            line["cpu_baseline"] = cpu
    return line


def allpairs_arm(ctx, steps=1):
    """BASELINE configs[3]: all-pairs top-k over one Yelp-shape document set (self join), rows sharded over the ranks
    (strong scaling: the job is fixed, every rank takes a block of rows; the only exchange is the final NCCL
    all-gather of (index, distance) per row).  Effective pairs/s = docs^2 / time."""
    from consistent__style_transfer_b200 import sharding
    a, torch, eng, dev, world, rank = ctx.a, ctx.torch, ctx.eng, ctx.dev, ctx.world, ctx.rank
    dist = ctx.dist
    ids, off, _, _ = workload.make_pairs(a.docs, a.shape, "independent", V=a.vocab, seed=1)
    ids = torch.from_numpy(ids).pin_memory().numpy(); off = torch.from_numpy(off).pin_memory().numpy()   # pinned host documents
    N, k = a.docs, a.topk
    blocks = sharding.row_blocks(N, world)
    r0, r1 = int(blocks[rank]), int(blocks[rank + 1])
    mx = int(np.diff(blocks).max())
    # warm-up on a small block: builds the word-distance table (one-off per embedding table) and the allocations
    t0 = time.perf_counter()
    eng.allpairs_topk(ids, off, ids, off, k, r0, min(r1, r0 + 256))
    ctx.barrier()
    t_warm = time.perf_counter() - t0
    t_idx = torch.zeros((mx, k), dtype=torch.int32, device=dev)
    t_dst = torch.zeros((mx, k), dtype=torch.float64, device=dev)
    g_idx = torch.empty((world * mx, k), dtype=torch.int32, device=dev) if world > 1 else None
    g_dst = torch.empty((world * mx, k), dtype=torch.float64, device=dev) if world > 1 else None
    infos, walls = [], []
    for _ in range(max(1, steps) + 1):                             # the first full pass sizes the workspace: not timed
        ctx.barrier()
        t0 = time.perf_counter()
        info = eng.allpairs_topk_cuda(ids, off, ids, off, k, r0, r1, out_idx=t_idx[:r1 - r0], out_dist=t_dst[:r1 - r0])
        if world > 1:                                              # results never left the device
            dist.all_gather_into_tensor(g_idx, t_idx); dist.all_gather_into_tensor(g_dst, t_dst)
        ctx.barrier()
        walls.append(time.perf_counter() - t0)
        infos.append(info)
    wall = ctx.max_over_ranks(statistics.median(walls[1:]))
    cnt = torch.tensor([infos[-1]["exact_round1"], infos[-1]["exact_round2"], infos[-1]["bounds"]], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(cnt)
    verified = None
    if rank == 0 and a.verify > 0:
        n = min(a.verify, N)
        rows = min(n, 512)
        sub_ids, sub_off = ids[:off[n]], off[:n + 1]
        vi, vd, _ = eng.allpairs_topk(sub_ids, sub_off, sub_ids, sub_off, k, 0, rows)
        # brute force of the same rows through the pair path of the engine (every pair solved exactly), one call
        lens = np.diff(sub_off)[:rows]
        rep_ids = np.concatenate([np.tile(sub_ids[sub_off[i]:sub_off[i + 1]], n) for i in range(rows)])
        rep_off = np.zeros(rows * n + 1, np.int64)
        np.cumsum(np.repeat(lens, n), out=rep_off[1:])
        all_ids = np.tile(sub_ids, rows)
        all_off = (np.tile(sub_off[:-1], rows) + np.repeat(np.arange(rows, dtype=np.int64) * int(sub_off[-1]), n))
        all_off = np.concatenate([all_off, [rows * int(sub_off[-1])]])
        d, _ = eng.wmd_pairs(rep_ids, rep_off, all_ids, all_off)
        d = d.reshape(rows, n)
        ok = True
        for i in range(rows):
            order = np.lexsort((np.arange(n), d[i]))[:k]
            ok = ok and np.array_equal(order.astype(np.int32), vi[i]) and d[i][order].tobytes() == vd[i].tobytes()
        verified = {"rows": int(rows), "block": int(n), "matches_bruteforce": bool(ok)}
    ctx.barrier()
    if rank != 0:
        return None
    return {"metric": "allpairs_effective_pairs_per_sec", "value": float(N) * N / wall, "unit": "pairs/s", "n_gpus": world,
            "steps": max(1, steps), "ms_per_step": 1e3 * wall, "higher_is_better": True, "scaling": "strong",
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"all-pairs top-{k} WMD, {N} x {N} {a.shape}-shape documents (self join), d={a.d}, V={a.vocab}",
                       "timing": "wall clock around the library call (H2D of the documents, bounds, exact solves, top-k left on the "
                                 "device) + the NCCL all-gather of the result, barrier on both sides, max over ranks; word-distance "
                                 f"table built in the warm-up ({t_warm:.2f}s incl. first allocations)"},
            "exact_solves": float(cnt[0] + cnt[1]), "exact_round1": float(cnt[0]), "exact_round2": float(cnt[1]),
            "bounds": float(cnt[2]), "pruned_fraction": 1.0 - float(cnt[0] + cnt[1]) / max(1.0, float(cnt[2])),
            "rank0_phase_ms": {kk: vv for kk, vv in infos[-1].items() if kk.startswith("ms_")},
            "verified": verified}


def sweep_arm(ctx, steps=3, lengths=None, mixed=True):
    """BASELINE configs[4]: fixed document lengths 8 -> 256 (both sides, independent draws), 2^18 pairs per GPU and length
    (2^14 from 128 tokens up), device-resident inputs, CUDA-event timed, max over ranks; the scores are all-gathered over
    NCCL inside the timed region when N > 1.  `mixed`: lengths uniform in [1, 256] per side -- no launch of the default
    path depends on the longest document of the batch, so its time should be the sum of its pairs' own costs."""
    a, torch, eng, dev, world, rank = ctx.a, ctx.torch, ctx.eng, ctx.dev, ctx.world, ctx.rank
    dist = ctx.dist
    lengths = lengths or [int(x) for x in a.lengths.split(",")]
    rows = {}

    def run(shape, n, seed, ml):
        ids1, off1, ids2, off2 = workload.make_pairs(n, shape, "independent", V=a.vocab, seed=seed + 1000 * rank)
        d = [torch.from_numpy(x).to(dev) for x in (ids1, off1, ids2, off2)]
        out = torch.empty(n, dtype=torch.float64, device=dev); st = torch.empty(n, dtype=torch.int32, device=dev)
        g_out = torch.empty(world * n, dtype=torch.float64, device=dev) if world > 1 else None

        def step():
            eng.wmd_pairs_cuda(d[0], d[1], d[2], d[3], ml, ml, out=out, status=st)
            if world > 1:
                dist.all_gather_into_tensor(g_out, out)
        for _ in range(3):
            step()
        ctx.barrier()
        ms = []
        for _ in range(max(3, steps)):
            ctx.flush.fill_(1)
            e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
            e0.record(); step(); e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        t = ctx.max_over_ranks(statistics.median(ms))
        stats = eng.last_stats()
        eng.set_serial(True); eng.set_profiling(True); eng.profile(reset=True)
        eng.wmd_pairs_cuda(d[0], d[1], d[2], d[3], ml, ml, out=out, status=st)
        torch.cuda.synchronize()
        prof = eng.profile(reset=True)
        eng.set_profiling(False); eng.set_serial(False)
        alg = 4 * stats["tokens"] + 4 * a.d * stats["uniques"] + 8 * n
        return {"pairs_per_gpu": n, "ms": t, "pairs_per_s": world * n / (t / 1e3),
                "mean_unique_tokens_per_side": stats["uniques"] / (2 * n),
                "algorithmic_gb_per_s_per_gpu": alg / (t / 1e3) / 1e9,
                "hbm_roofline_frac": alg / (t / 1e3) / 1e9 / hbm_peak()[0],
                "table_mode_gb_per_s_per_gpu": (4 * stats["tokens"] + 4 * stats["cells"] + 8 * n) / (t / 1e3) / 1e9,
                "kernel_ms_serial": {k: v["ms"] for k, v in prof.items() if v["launches"] > 0}}, (off1, off2)

    for L in lengths:
        rows[str(L)], _ = run(f"fixed:{L}", 1 << (18 if L < 128 else 14), L, L)
    mixed_rec = None
    if mixed and len(lengths) >= 2:
        n = 1 << 16
        rec, (off1, off2) = run("uniform:1-256", n, 77, 256)
        # the pairs' own costs: per-pair time of the fixed-length runs, interpolated log-log at sqrt(n1 * n2)
        Ls = np.array(sorted(int(k) for k in rows), dtype=np.float64)
        per_pair = np.array([rows[str(int(L))]["ms"] / rows[str(int(L))]["pairs_per_gpu"] for L in Ls])
        n1, n2 = np.maximum(np.diff(off1), 1).astype(np.float64), np.maximum(np.diff(off2), 1).astype(np.float64)
        predict = lambda eff: float(np.exp(np.interp(np.log(eff), np.log(Ls), np.log(per_pair))).sum())
        pred_geo, pred_max = predict(np.sqrt(n1 * n2)), predict(np.maximum(n1, n2))
        rec["predicted_ms_from_fixed_length_rates"] = {"at_sqrt_n1_n2": pred_geo, "at_max_n1_n2": pred_max}
        rec["measured_over_predicted"] = {"at_sqrt_n1_n2": rec["ms"] / pred_geo, "at_max_n1_n2": rec["ms"] / pred_max}
        rec["note"] = ("a pair of n1 x n2 tokens is charged the per-pair time of the fixed-length run at sqrt(n1 n2) (same number of "
                       "cost cells) or at max(n1, n2) (the solver's searches grow with the longer side); no launch of the default path "
                       "is sized by the longest document of the batch, so the mixed batch costs the sum of its pairs' own costs")
        mixed_rec = rec
    if rank != 0:
        return None
    return {"metric": "wmd_length_sweep_pairs_per_sec", "unit": "pairs/s", "n_gpus": world, "dtype": "f64", "data": "synthetic",
            "scaling": "weak",
            "config": {"workload": f"fixed lengths {','.join(map(str, lengths))} both sides, independent, d={a.d}, V={a.vocab}, 2^18 pairs per GPU (2^14 from 128 tokens)",
                       "l2": "256 MiB flush write between timed steps",
                       "timing": "median of the timed steps per rank (CUDA events), max over ranks",
                       "roofline": "hbm_roofline_frac charges SURVEY 8(d)'s algorithmic bytes (4(n1+n2) + 4d(u1+u2) + 8 per pair); the default "
                                   "path takes 4 B per cost cell from the word-distance table instead of d floats per row, so at short "
                                   "lengths it finishes faster than HBM could deliver the convention's bytes (fraction > 1) -- "
                                   "table_mode_gb_per_s_per_gpu is what it really has to fetch; from 32 tokens up the exact solver binds"},
            "peak_hbm_gbs": hbm_peak()[0], "lengths": rows,
            "mixed_uniform_1_256": mixed_rec}


def latency_arm(a):
    """BASELINE configs[2] through the reference's own call surface, one pretrain collate batch per call (src/loader.py:60:
    256 Yelp pairs or 128 book pairs of noised sentences, d = 100): `WMDdistance.cal_wmd_label(lists, lists, tokenizer)`,
    `collate_pretrain(...)(samples)` and `calculate_wmd_scores(strings, strings, model)`, wall clock per call, next to the
    oracle port of the reference loop on ONE core (the reference's collate runs single-threaded in the main process)."""
    import random

    from consistent__style_transfer_b200 import content_preserve as cp
    from consistent__style_transfer_b200.loader import LabelPrefetcher, collate_pretrain
    from consistent__style_transfer_b200.wmd import WMDdistance
    from oracle import wmd_oracle
    V = a.vocab
    table = workload.make_table(V, 100, seed=0)                    # d = 100: the reference's real embedding width
    words = ["w%06d" % i for i in range(V)]

    class Tok:                                                     # shape of src/vocab.py:BPETokenizer as the path uses it
        tokenizer = None

        def __init__(self):
            self.tokenizer = self

        def id_to_token(self, i):
            return words[i - 4] if 4 <= i < V + 4 else None

        def ids_to_tokens(self, ids):
            return [self.id_to_token(i) for i in ids]

        def __len__(self):
            return V + 4

    tok = Tok()
    w = WMDdistance.from_embeddings(words, table, normalize=False)
    kv = wmd_oracle.KeyedVectorsOracle(words, table)
    ow = wmd_oracle.WMDdistanceOracle(kv)
    out = {}
    for shape, B in (("yelp", 256), ("book", 128)):
        ids1, off1, ids2, off2 = workload.make_pairs(B * 32, shape, "noised", V=V, seed=5, batch=B)
        lists = lambda ids, off, lo, hi: [(ids[off[p]:off[p + 1]] + 4).tolist() for p in range(lo, hi)]
        calls = [(lists(ids1, off1, b * B, (b + 1) * B), lists(ids2, off2, b * B, (b + 1) * B)) for b in range(32)]
        samples = [[(s, i % 2) for i, s in enumerate(c[0])] for c in calls]

        def med(f, args_list, reps=3):
            for x in args_list[:4]:
                f(x)
            lat = []
            for _ in range(reps):
                for x in args_list:
                    t0 = time.perf_counter(); f(x); lat.append(time.perf_counter() - t0)
            return float(np.median(lat) * 1e6), float(np.quantile(lat, 0.99) * 1e6)

        label_us, label_p99 = med(lambda c: w.cal_wmd_label(c[0], c[1], tok), calls)
        np.random.seed(1); random.seed(2)
        coll = collate_pretrain(tok, w)
        collate_us, _ = med(coll, samples)
        t0 = time.perf_counter()
        nb = sum(1 for _ in LabelPrefetcher(samples * 3, tok, w))
        prefetch_us = (time.perf_counter() - t0) / nb * 1e6
        strs = [([" ".join(words[t - 4] for t in s) for s in c[0]], [" ".join(words[t - 4] for t in s) for s in c[1]]) for c in calls[:8]]
        score_us, _ = med(lambda c: cp.calculate_wmd_scores(c[0], c[1], w.model), strs)
        c = calls[0]
        t0 = time.perf_counter()
        ow.cal_wmd_label(c[0], c[1], tok)
        cpu_s = time.perf_counter() - t0
        out[f"{shape}_batch{B}"] = {
            "cal_wmd_label_us": label_us, "cal_wmd_label_p99_us": label_p99, "collate_pretrain_us": collate_us,
            "collate_with_label_prefetch_us": prefetch_us, "calculate_wmd_scores_us": score_us,
            "pairs_per_s_through_cal_wmd_label": float(B / (label_us * 1e-6)),
            "cpu_reference_port_1core_ms": cpu_s * 1e3, "speedup_vs_1core": float(cpu_s / (label_us * 1e-6))}
    w.model.wv.close()
    return {"metric": "wmd_batch_latency", "unit": "us", "n_gpus": 1, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "one pretrain collate batch per call (noised pairs, d=100, V=%d): python lists of tokenizer ids / "
                                   "strings in, python floats / tensors out; median wall time per call" % V},
            "batches": out}


def run_b200(a):
    ctx = Ctx(a)
    cpu = None
    if a.mode == "pairs" and ctx.rank == 0 and ctx.world == 1 and not a.no_cpu_baseline:
        # CPU baseline first (fork-based workers must start before CUDA is initialised)
        cores = os.cpu_count() or 1
        table = workload.make_table(a.vocab, a.d, seed=0)
        pairs = workload.make_pairs(max(a.cpu_sample or 3000 * cores, 40000), a.shape, a.variant, V=a.vocab, seed=1)
        n_sample = a.cpu_sample or 3000 * cores
        pool = CpuPool(table, pairs, cores)
        wall = pool.run(0, n_sample)
        pool.close()
        c_port = cpu_c_port(table, pairs, 40000, cores)
        one = cpu_one_core(table, pairs, 1500)
        cpu = {"value": n_sample / wall, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"first {n_sample} pairs of the same workload split over {cores} worker processes: "
                         f"gensim-shaped python per-pair loop (numpy float32 per cell) + C emd_hat; {wall:.1f}s wall",
               "one_core_value": one,
               "one_core_note": "the same loop in ONE process on 1 500 pairs: the reference's real configuration "
                                "(collate runs in the main process, src/main_pretrain.py:120-122)",
               "compiled_c_port_value": c_port,
               "compiled_c_port_note": f"all-C oracle (oracle/wmd_oracle.c) on 40000 pairs, {cores} pthreads"}
    if a.mode == "latency":
        print(json.dumps(latency_arm(a)), flush=True)
        return
    ctx.start_gpu()
    line = None
    if a.mode == "pairs":
        line = pairs_arm(ctx, cpu)
        if not a.no_extras:
            extras = {}
            for name, fn in (("allpairs", lambda: allpairs_arm(ctx, steps=2)), ("sweep", lambda: sweep_arm(ctx, steps=3))):
                try:
                    extras[name] = fn()
                except Exception as exc:                                   # a sub-record must never cost the headline
                    if ctx.world > 1:
                        raise
                    extras[name] = {"error": str(exc)[:300]}
            if ctx.rank == 0:
                line.update(extras)
    elif a.mode == "allpairs":
        line = allpairs_arm(ctx, steps=a.steps)
    elif a.mode == "sweep":
        line = sweep_arm(ctx, steps=a.steps)
    ctx.close()
    if a.mode == "pairs" and not a.no_extras and ctx.rank == 0 and ctx.world == 1:
        try:
            line["latency"] = latency_arm(a)
        except Exception as exc:
            line["latency"] = {"error": str(exc)[:300]}
    if ctx.rank == 0 and line is not None:
        print(json.dumps(line), flush=True)


if __name__ == "__main__":
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
