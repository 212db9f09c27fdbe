#!/usr/bin/env python
"""Randomised parity stress of the pair path against the C oracle (run on the GPU box):
    python tools/stress_parity.py [pairs_per_config]
Small vocabularies (heavy token sharing: cancellation, ties, degenerate transportation problems), tiny and
long documents, several embedding widths, rank tables and OOV ids.  Every value must match the oracle bit for bit."""
import sys
import time

import numpy as np

sys.path.insert(0, ".")
from consistent__style_transfer_b200 import workload  # noqa: E402
from consistent__style_transfer_b200.engine import WMDEngine  # noqa: E402
from oracle import wmd_oracle  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 200_000
CONFIGS = [
    # shape, variant, V, d, use_rank, oov_rate
    ("yelp", "independent", 10_000, 300, False, 0.0),
    ("yelp", "noised", 10_000, 300, False, 0.0),
    ("yelp", "independent", 40, 300, True, 0.05),
    ("yelp", "noised", 25, 100, False, 0.0),
    ("book", "independent", 2_000, 100, True, 0.02),
    ("book", "noised", 60, 64, False, 0.0),
    ("fixed:1", "independent", 30, 16, False, 0.1),
    ("fixed:2", "independent", 12, 8, False, 0.0),
    ("fixed:3", "independent", 8, 5, True, 0.0),
    ("fixed:33", "independent", 500, 40, False, 0.0),
    ("fixed:70", "independent", 3_000, 24, False, 0.0),
    ("fixed:200", "independent", 5_000, 12, False, 0.0),
    ("uniform:1-100", "independent", 200, 300, False, 0.01),
    ("uniform:20-200", "noised", 600, 20, True, 0.0),
    ("uniform:33-64", "independent", 64, 48, False, 0.0),
    # every wide solver class, both orientations, the 257-row case (solve_wide.cuh)
    ("uniform:1-256", "independent", 10_000, 300, False, 0.0),
    ("fixed:256", "independent", 700, 16, True, 0.0),
    ("uniform:100-200", "noised", 3_000, 64, False, 0.01),
]
bad = 0
for shape, variant, V, d, use_rank, oov in CONFIGS:
    n = max(2000, N // 50) if shape.startswith("uniform:") or (shape.startswith("fixed:") and int(shape.split(":")[1]) >= 30) else N
    rng = np.random.default_rng(hash((shape, variant, V, d)) % (2 ** 32))
    table = workload.make_table(V, d, seed=int(rng.integers(1 << 30)))
    ids1, off1, ids2, off2 = workload.make_pairs(n, shape, variant, V=V, seed=int(rng.integers(1 << 30)))
    ids1 = ids1.copy(); ids2 = ids2.copy()
    if oov > 0:
        ids1[rng.random(len(ids1)) < oov] = -1
        ids2[rng.random(len(ids2)) < oov] = -1
    rank = rng.permutation(V).astype(np.int32) if use_rank else None
    eng = WMDEngine(table, rank=rank)
    t0 = time.perf_counter()
    got, st = eng.wmd_pairs(ids1, off1, ids2, off2)
    t1 = time.perf_counter()
    want, wst = wmd_oracle.batch_wmd(table, ids1, off1, ids2, off2, rank=rank, nthreads=16)
    t2 = time.perf_counter()
    same = got.tobytes() == want.tobytes() and np.array_equal(st, wst)
    nbad = int(np.sum((got != want) & ~(np.isnan(got) & np.isnan(want)))) + int(np.sum(st != wst))
    bad += nbad
    print(f"{shape:10s} {variant:12s} V={V:6d} d={d:4d} rank={use_rank!s:5} oov={oov:4.2f} pairs={n:8d}: "
          f"{'OK ' if same else 'MISMATCH'} bad={nbad} status={np.bincount(st, minlength=4).tolist()} gpu {t1 - t0:.2f}s oracle {t2 - t1:.2f}s",
          flush=True)
    eng.close()
print("TOTAL MISMATCHES", bad)
sys.exit(1 if bad else 0)
