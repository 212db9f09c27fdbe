#!/usr/bin/env python
"""Python model of the class-A transport solver's search (solve.cuh: transport_solve_small), used to rank
algorithmic variants offline by the quantities that cost instructions on the GPU: searches, column selections,
row relaxations and multi-hop augmentations per pair.  On the bench workload the model's counts equal the
per-SASS-instruction execution counts ncu reports for the kernel (70.3 selections / 39.1 relaxations per pair
before the orientation change), so a variant can be judged without GPU time.

    python tools/solver_model.py [pairs]
    python tools/solver_model.py --long [pairs]      # long documents: fixed 256 tokens against mixed lengths 1 .. 256

Variants reported: rows = supplying side (the original), rows = the side with more nodes (built), the latter
with searches continuing after intact multi-hop augmentations (built), and on top of that the reduced-cost start with
its greedy pass (built in round 2).  Every variant must return the same optimum.
"""
import sys

import numpy as np

sys.path.insert(0, ".")
from consistent__style_transfer_b200 import workload  # noqa: E402


def solve(C, sup, dem, cont, reduced_start=False):
    """Successive shortest paths, row by row, Dijkstra over the columns; a direct arc out of the root never ends a
    search; with cont a multi-hop augmentation that empties no reverse arc does not either.  reduced_start: duals from a
    row and a column reduction, then one greedy pass shipping every row along a tight arc into a column with a deficit
    (what transport_solve_small does before its first search)."""
    m, n = C.shape
    u = np.zeros(m, np.int64); v = np.zeros(n, np.int64)
    F = np.zeros((m, n), np.int64); dem = dem.copy(); sup = sup.copy()
    nsearch = nsel = nrelax = naug = 0
    if reduced_start:
        u = C.min(1).astype(np.int64)
        v = (C - u[:, None]).min(0).astype(np.int64)
        for r in range(m):
            tight = np.nonzero((C[r] - u[r] - v == 0) & (dem > 0))[0]
            if len(tight) and sup[r] > 0:
                j = int(tight[0]); amt = min(sup[r], dem[j]); F[r, j] += amt; sup[r] -= amt; dem[j] -= amt
    for r in range(m):
        while sup[r] > 0:
            nsearch += 1
            used = np.zeros(n, bool); tree = np.zeros(m, bool); tree[r] = True
            rdist = np.zeros(m, np.int64); rpred = -np.ones(m, np.int64)
            minv = C[r] - u[r] - v; way = np.full(n, r)
            while True:
                key = np.where(used, 1 << 60, minv); j = int(key.argmin()); delta = key[j]; used[j] = True; nsel += 1
                if dem[j] > 0:
                    if way[j] == r:
                        amt = min(sup[r], dem[j]); F[r, j] += amt; dem[j] -= amt; sup[r] -= amt
                        if sup[r] == 0:
                            break
                    else:
                        path = []; jj = j
                        while True:
                            i = way[jj]; path.append((i, jj))
                            if i == r:
                                break
                            jj = rpred[i]
                        amt = min(sup[r], dem[j])
                        for i, jj in path:
                            if i != r:
                                amt = min(amt, F[i, rpred[i]])
                        emptied = False
                        for i, jj in path:
                            F[i, jj] += amt
                            if i != r:
                                F[i, rpred[i]] -= amt
                                emptied |= F[i, rpred[i]] == 0
                        sup[r] -= amt; dem[j] -= amt; naug += 1
                        if not cont or emptied or sup[r] == 0 or dem[j] > 0:
                            break
                for i in np.nonzero((F[:, j] > 0) & ~tree)[0]:
                    tree[i] = True; rdist[i] = delta; rpred[i] = j; nrelax += 1
                    cand = delta + C[i] - u[i] - v
                    upd = (cand < minv) & ~used
                    minv[upd] = cand[upd]; way[upd] = i
            u[tree] += delta - rdist[tree]; v[used] -= delta - minv[used]
    return int((F * C).sum()), np.array([nsearch, nsel, nrelax, naug])


def residual_problem(T, a, b):
    """The balanced integer problem pyemd solves for documents a, b: rows = supplying side."""
    ua, ca = np.unique(a, return_counts=True); ub, cb = np.unique(b, return_counts=True)
    allw = np.union1d(ua, ub)
    P = np.zeros(len(allw)); Q = np.zeros(len(allw))
    P[np.searchsorted(allw, ua)] = ca / len(a); Q[np.searchsorted(allw, ub)] = cb / len(b)
    X = T[allw]
    D = np.sqrt(((X[:, None, :] - X[None, :, :]) ** 2).sum(-1, dtype=np.float32)).astype(np.float64)
    inA = np.isin(allw, ua); inB = np.isin(allw, ub)
    Dm = np.where(inA[:, None] & inB[None, :], D, 0); Dm = np.maximum(Dm, Dm.T); np.fill_diagonal(Dm, 0)
    if Dm.max() == 0:
        return None
    Pc = np.where(P < Q, 0, P - Q); Qc = np.where(P < Q, Q - P, 0)
    PQn = 1e6 / max(P.sum(), Q.sum())
    iP = np.floor(Pc * PQn + 0.5).astype(np.int64); iQ = np.floor(Qc * PQn + 0.5).astype(np.int64)
    iC = np.floor(Dm * (1e6 / Dm.max()) + 0.5).astype(np.int64)
    if iP.sum() < iQ.sum():
        iP, iQ = iQ, iP; iC = iC.T
    r = np.nonzero(iP > 0)[0]; c = np.nonzero(iQ > 0)[0]
    if len(r) == 0 or len(c) == 0:
        return None
    C = iC[np.ix_(r, c)]; s = iP[r]; t = iQ[c]; diff = s.sum() - t.sum()
    if diff > 0:
        C = np.hstack([C, np.zeros((len(r), 1), np.int64)]); t = np.append(t, diff)
    return C, s, t


def long_documents(npairs):
    """What the wide solver (solve_wide.cuh) sees: column selections / row relaxations per residual problem, rows = the
    side with more nodes, for 256-token pairs (near-equal masses: almost assignment problems) and for pairs of mixed
    lengths (masses 1/n1 against 1/n2: every node needs several partners) -- the reason a mixed batch costs more than
    the fixed-length runs predict (profiles/README.md)."""
    T = workload.make_table(10_000, 300, seed=0)
    for shape in ("fixed:256", "uniform:1-256"):
        ids1, off1, ids2, off2 = workload.make_pairs(npairs, shape, "independent", V=10_000, seed=1)
        print(shape)
        for p in range(npairs):
            pr = residual_problem(T, ids1[off1[p]:off1[p + 1]], ids2[off2[p]:off2[p + 1]])
            if pr is None:
                continue
            C, s, t = pr
            if C.shape[0] < C.shape[1]:
                C, s, t = C.T.copy(), t.copy(), s.copy()
            if C.shape[0] <= 32:
                continue
            _, cnt = solve(C, s, t, True)
            print(f"  {C.shape[0]:3d} x {C.shape[1]:3d}: searches {cnt[0]:5d}, column selections {cnt[1]:6d}, row relaxations {cnt[2]:6d}")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "--long":
        long_documents(int(sys.argv[2]) if len(sys.argv) > 2 else 8)
        return
    npairs = int(sys.argv[1]) if len(sys.argv) > 1 else 400
    T = workload.make_table(10_000, 300, seed=0)
    ids1, off1, ids2, off2 = workload.make_pairs(npairs, "yelp", "independent", V=10_000, seed=1)
    tot = {}; n = 0
    for p in range(npairs):
        pr = residual_problem(T, ids1[off1[p]:off1[p + 1]], ids2[off2[p]:off2[p + 1]])
        if pr is None:
            continue
        C, s, t = pr; n += 1
        big = (C, s, t) if C.shape[0] >= C.shape[1] else (C.T.copy(), t.copy(), s.copy())
        ref = None
        for name, (prob, cont, red) in {"rows = supplying side": ((C, s, t), False, False), "rows = larger side": (big, False, False),
                                        "rows = larger side, searches continue": (big, True, False),
                                        "... and a reduced-cost start with a greedy pass": (big, True, True)}.items():
            val, cnt = solve(*prob, cont, red)
            ref = val if ref is None else ref
            assert val == ref, (name, val, ref)
            tot[name] = tot.get(name, 0) + cnt
    print(f"{n} pairs; per pair: searches, column selections, row relaxations, multi-hop augmentations")
    for name, cnt in tot.items():
        print(f"  {name:40s} {np.round(cnt / n, 1).tolist()}")


if __name__ == "__main__":
    main()
