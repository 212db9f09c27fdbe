"""GPU parity tests: libwmd_b200.so (through its C ABI) against the CPU oracle on the same inputs.

Bars (BASELINE.json north_star): WMD within 1e-6 relative of the reference CPU WMD (here the
expectation is tighter: the integer optimum is unique, so values agree to the last bit unless a
double rounding differs -- we assert <= 1e-12 relative and report exact-equality counts);
nBOW counts / weights, statuses and RWMD argmins bit-exact.
"""
import math

import numpy as np
import pytest

from consistent__style_transfer_b200 import workload

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng_mod():
    from consistent__style_transfer_b200 import engine
    return engine


@pytest.fixture(params=[True, False], ids=["table", "direct"])
def dtab(request):
    """Both cost sources of the pair entries: the word-distance table with its fused warp-per-pair kernel (the
    default policy) and the direct path that recomputes every distance from the embedding rows."""
    return request.param


def _assert_wmd_equal(got, st, want, wst, rtol=1e-12):
    assert np.array_equal(st, wst), np.nonzero(st != wst)[0][:10]
    fin = np.isfinite(want)
    assert np.array_equal(np.isfinite(got), fin)
    assert np.array_equal(got[~fin], want[~fin])              # +inf where the oracle says +inf
    np.testing.assert_allclose(got[fin], want[fin], rtol=rtol, atol=0.0)


@pytest.mark.parametrize("d", [5, 100, 130, 300])
def test_normalize_rows_bit_exact(eng_mod, oracle, d):
    rng = np.random.default_rng(d)
    raw = rng.standard_normal((257, d)).astype(np.float32) * 3.0
    e = eng_mod.WMDEngine(raw, normalize=True)
    want = oracle.init_sims_replace(raw)
    assert e.table().tobytes() == want.tobytes()
    e.close()


@pytest.mark.parametrize("use_rank", [False, True])
def test_nbow_bit_exact(eng_mod, oracle, use_rank):
    V = 500
    rng = np.random.default_rng(3)
    table = workload.make_table(V, 16, seed=1)
    rank = rng.permutation(V).astype(np.int32) if use_rank else None
    ids, off, _, _ = workload.make_pairs(300, "book", "independent", V=V, seed=4)
    ids = ids.copy(); ids[rng.random(len(ids)) < 0.1] = -1
    e = eng_mod.WMDEngine(table, rank=rank)
    rows, counts, weights, uniq = e.nbow(ids, off)
    for p in range(300):
        r, c, w = oracle.nbow(ids[off[p]:off[p + 1]], rank)
        u = uniq[p]
        assert u == len(r)
        a = off[p]
        assert np.array_equal(rows[a:a + u], r) and np.array_equal(counts[a:a + u], c)
        assert weights[a:a + u].tobytes() == w.tobytes()
    e.close()


@pytest.mark.parametrize("d,shape,variant,B", [
    (100, "yelp", "noised", 4000), (300, "yelp", "independent", 4000), (300, "yelp", "noised", 2000),
    (100, "book", "noised", 1500), (300, "book", "independent", 1000), (7, "yelp", "independent", 500),
    (130, "yelp", "independent", 500),
    # widths whose numpy leaves start 32 floats apart (bank-aligned: the cost kernel walks the leaves one after the other),
    # a width with 8 leaves, and one that needs the two-leaf variant
    (256, "yelp", "independent", 800), (512, "yelp", "independent", 800), (1000, "yelp", "noised", 300),
    (200, "book", "independent", 400),
])
def test_wmd_pairs_match_oracle(eng_mod, oracle, d, shape, variant, B, dtab):
    V = 2000
    table = workload.make_table(V, d, seed=2)
    ids1, off1, ids2, off2 = workload.make_pairs(B, shape, variant, V=V, seed=7)
    rng = np.random.default_rng(5)
    ids1 = ids1.copy(); ids1[rng.random(len(ids1)) < 0.03] = -1      # OOV tokens
    e = eng_mod.WMDEngine(table, distance_table=dtab)
    got, st = e.wmd_pairs(ids1, off1, ids2, off2)
    want, wst = oracle.batch_wmd(table, ids1, off1, ids2, off2, nthreads=8)
    _assert_wmd_equal(got, st, want, wst)
    e.close()


def test_wmd_pairs_with_rank_and_token_map(eng_mod, oracle, dtab):
    V = 800
    rng = np.random.default_rng(12)
    table = workload.make_table(V, 100, seed=3)
    rank = rng.permutation(V).astype(np.int32)
    ntok = 1000                                                   # tokenizer ids 0..999, some unmapped
    tmap = np.full(ntok, -1, np.int32)
    tmap[rng.permutation(ntok)[:V]] = np.arange(V, dtype=np.int32)
    t1, off1, t2, off2 = workload.make_pairs(1500, "yelp", "noised", V=ntok, seed=9)
    e = eng_mod.WMDEngine(table, rank=rank, token_map=tmap, distance_table=dtab)
    got, st = e.wmd_pairs(t1, off1, t2, off2)
    want, wst = oracle.batch_wmd(table, tmap[t1], off1, tmap[t2], off2, rank=rank, nthreads=8)
    _assert_wmd_equal(got, st, want, wst)
    e.close()


def test_early_outs_and_edge_cases(eng_mod, oracle, dtab):
    table = workload.make_table(50, 16, seed=3)
    table[7] = table[6]
    docs1 = [[-1], [1], [1, 1], [6], [1, 2], [], [3], [5, 5, 5, 9], [99999]]
    docs2 = [[1], [], [1], [7], [2, 1], [], [9], [9, 5], [1]]
    ids1, off1 = workload.to_csr(docs1); ids2, off2 = workload.to_csr(docs2)
    e = eng_mod.WMDEngine(table, distance_table=dtab)
    got, st = e.wmd_pairs(ids1, off1, ids2, off2)
    want, wst = oracle.batch_wmd(table, np.where(ids1 >= 50, -1, ids1).astype(np.int32), off1, ids2, off2)
    assert list(st) == [1, 1, 2, 3, 0, 1, 0, 0, 1]
    _assert_wmd_equal(got, st, want, wst)
    assert got[4] == 0.0
    # empty batch
    z = np.zeros(1, np.int64)
    g0, s0 = e.wmd_pairs(np.zeros(0, np.int32), z, np.zeros(0, np.int32), z)
    assert g0.shape == (0,) and s0.shape == (0,)
    # too-long document is an error, not a silent truncation
    long_ids, long_off = workload.to_csr([[1] * 300])
    with pytest.raises(RuntimeError):
        e.wmd_pairs(long_ids, long_off, long_ids, long_off)
    e.close()


@pytest.mark.parametrize("L", [8, 33, 64, 128, 256])
def test_length_sweep_matches_oracle(eng_mod, oracle, L, dtab):
    V = 10000
    table = workload.make_table(V, 300, seed=0)
    B = 96 if L >= 128 else 400
    ids1, off1, ids2, off2 = workload.make_pairs(B, f"fixed:{L}", "independent", V=V, seed=L)
    e = eng_mod.WMDEngine(table, distance_table=dtab)
    got, st = e.wmd_pairs(ids1, off1, ids2, off2)
    want, wst = oracle.batch_wmd(table, ids1, off1, ids2, off2, nthreads=8)
    _assert_wmd_equal(got, st, want, wst)
    e.close()


@pytest.mark.parametrize("shape,variant,V,d,B", [
    # long documents over small vocabularies: unequal masses, partial cancellation, tied costs, residual problems
    # of every solver class (A and the wide instances) and packed as well as block-split cost stages in one launch
    ("uniform:1-90", "independent", 3000, 300, 1500), ("uniform:1-256", "independent", 400, 64, 600), ("fixed:256", "independent", 300, 8, 200),
    ("uniform:30-70", "noised", 150, 100, 800), ("fixed:100", "independent", 120, 32, 500),
    ("fixed:200", "noised", 5000, 16, 300), ("uniform:40-45", "independent", 10000, 300, 1000),
    # around the fused kernel's limits: 32-token documents whose residual problem needs 33 columns, documents either side of 32 tokens
    ("uniform:28-36", "independent", 10000, 64, 1500), ("fixed:32", "independent", 10000, 32, 600), ("uniform:30-33", "noised", 4000, 100, 900),
])
def test_mixed_long_documents_match_oracle(eng_mod, oracle, shape, variant, V, d, B, dtab):
    table = workload.make_table(V, d, seed=11)
    ids1, off1, ids2, off2 = workload.make_pairs(B, shape, variant, V=V, seed=13)
    e = eng_mod.WMDEngine(table, distance_table=dtab)
    got, st = e.wmd_pairs(ids1, off1, ids2, off2)
    want, wst = oracle.batch_wmd(table, ids1, off1, ids2, off2, nthreads=8)
    _assert_wmd_equal(got, st, want, wst)
    e.close()


def _chain_table(n, eps=0.002, h=0.2, d=8):
    """Rows 0..n on a short arc of the unit circle, rows n+1..2n+1 the same arc lifted to latitude h: document
    {1..n} against {n+1..2n} (lifted points 0..n-1) assigns point i to lifted point i until the last row finds only
    lifted point 0 free, and the cheapest way there shifts every earlier assignment by one: an augmenting path of n hops."""
    T = np.zeros((2 * n + 2, d), np.float32)
    for i in range(n + 1):
        T[i, :3] = [np.cos(i * eps), np.sin(i * eps), 0.0]
        T[n + 1 + i, :3] = [np.cos(i * eps) * np.cos(h), np.sin(i * eps) * np.cos(h), np.sin(h)]
    return T


@pytest.mark.parametrize("n", [12, 31, 33, 64, 100, 200, 256])
def test_long_augmenting_paths_match_oracle(eng_mod, oracle, n, dtab):
    # paths longer than 32 hops are walked in pieces by the wide solver (solve_wide.cuh: transport_solve_wide)
    table = _chain_table(n)
    fwd = (np.arange(1, n + 1), n + 1 + np.arange(0, n))
    docs1, docs2 = [], []
    rng = np.random.default_rng(n)
    for k in range(24):
        a, b = fwd if k % 2 == 0 else fwd[::-1]
        if k >= 2:                                                # variations: repeated tokens (unequal masses), dropped tokens
            a = np.concatenate([a, rng.choice(a, size=rng.integers(0, 4))]) if len(a) < 250 else a
            b = np.delete(b, rng.integers(0, len(b), size=rng.integers(0, 3)))
        docs1.append(a.astype(np.int32)); docs2.append(b.astype(np.int32))
    ids1, off1 = workload.to_csr(docs1)
    ids2, off2 = workload.to_csr(docs2)
    e = eng_mod.WMDEngine(table, distance_table=dtab)
    got, st = e.wmd_pairs(ids1, off1, ids2, off2)
    want, wst = oracle.batch_wmd(table, ids1, off1, ids2, off2, nthreads=4)
    _assert_wmd_equal(got, st, want, wst)
    assert np.array_equal(got, want)
    e.close()


def test_wide_solver_shapes_match_oracle(eng_mod, oracle, dtab):
    # solve_wide.cuh: every instance (column words of the shorter side 1 .. 8), both orientations (the supplying side the
    # longer or the shorter one), the 257-row case (256 nodes a side plus the dummy), heavy cancellation (nodes without
    # residual mass still count for maxC)
    V = 700
    table = workload.make_table(V, 12, seed=5)
    rng = np.random.default_rng(17)
    docs1, docs2 = [], []
    for long_n in (33, 64, 65, 100, 129, 161, 193, 225, 256):
        for short_n in (1, 9, 31, 33, 64, 97, 128, 160, 192, 224, 256):
            if short_n > long_n:
                continue
            a = rng.choice(V, size=long_n, replace=False)
            b = rng.choice(V, size=short_n, replace=False)
            if (long_n + short_n) % 3 == 0:                      # shared tokens: cancellation on both sides
                k = min(short_n, long_n) // 2
                b[:k] = a[:k]
            if (long_n + short_n) % 4 == 1:                      # repeated tokens: unequal masses
                b = np.concatenate([b, b[:min(len(b), 256 - len(b))][:7]])
            for x, y in ((a, b), (b, a)):
                docs1.append(x.astype(np.int32)); docs2.append(y.astype(np.int32))
    full = rng.permutation(V)[:512]
    docs1.append(full[:256].astype(np.int32)); docs2.append(full[256:].astype(np.int32))                 # 256 x 256, equal masses
    docs1.append(np.concatenate([full[:255], full[:1]]).astype(np.int32)); docs2.append(full[256:].astype(np.int32))      # 255 x 256 + dummy
    docs1.append(full[:256].astype(np.int32)); docs2.append(np.concatenate([full[256:511], full[256:257]]).astype(np.int32))
    docs1.append(full[:256].astype(np.int32)); docs2.append(full[100:356].astype(np.int32))              # 156 shared tokens
    ids1, off1 = workload.to_csr(docs1)
    ids2, off2 = workload.to_csr(docs2)
    e = eng_mod.WMDEngine(table, distance_table=dtab)
    got, st = e.wmd_pairs(ids1, off1, ids2, off2)
    want, wst = oracle.batch_wmd(table, ids1, off1, ids2, off2, nthreads=8)
    _assert_wmd_equal(got, st, want, wst)
    assert np.array_equal(got, want)
    e.close()


@pytest.mark.parametrize("shape,variant,V,d,B,rank", [
    ("yelp", "independent", 3000, 300, 3000, False), ("yelp", "noised", 500, 100, 3000, True),
    ("book", "independent", 2000, 64, 1000, False), ("uniform:1-120", "independent", 700, 300, 400, False),
    ("fixed:256", "independent", 900, 24, 60, True),
])
def test_distance_table_path_is_bit_identical(eng_mod, oracle, shape, variant, V, d, B, rank):
    # wmd_set_distance_table: tiles gathered from the V x V table instead of recomputed; same bits as the direct path and the oracle
    table = workload.make_table(V, d, seed=21)
    ids1, off1, ids2, off2 = workload.make_pairs(B, shape, variant, V=V, seed=23)
    rng = np.random.default_rng(3)
    ids2 = ids2.copy(); ids2[rng.random(len(ids2)) < 0.02] = -1
    rk = rng.permutation(V).astype(np.int32) if rank else None
    e = eng_mod.WMDEngine(table, rank=rk, distance_table=False)
    assert not e.distance_table_info()["resident"]
    direct, st0 = e.wmd_pairs(ids1, off1, ids2, off2)
    e.set_distance_table(True)
    assert e.distance_table_info()["resident"]
    got, st = e.wmd_pairs(ids1, off1, ids2, off2)
    lb_t = e.rwmd_pairs(ids1, off1, ids2, off2)
    e.set_distance_table(False)
    again, st2 = e.wmd_pairs(ids1, off1, ids2, off2)
    lb_d = e.rwmd_pairs(ids1, off1, ids2, off2)
    assert got.tobytes() == direct.tobytes() == again.tobytes()
    assert np.array_equal(st, st0) and np.array_equal(st2, st0)
    for k in lb_t:
        assert np.asarray(lb_t[k]).tobytes() == np.asarray(lb_d[k]).tobytes(), k
    want, wst = oracle.batch_wmd(table, ids1, off1, ids2, off2, rank=rk, nthreads=8)
    _assert_wmd_equal(got, st, want, wst)
    e.close()


def test_rwmd_matches_oracle(eng_mod, oracle):
    V = 600
    table = workload.make_table(V, 100, seed=6)
    ids1, off1, ids2, off2 = workload.make_pairs(400, "book", "independent", V=V, seed=13)
    e = eng_mod.WMDEngine(table)
    r = e.rwmd_pairs(ids1, off1, ids2, off2)
    wmd, _ = e.wmd_pairs(ids1, off1, ids2, off2)
    for p in range(400):
        w = oracle.rwmd_pair(table, ids1[off1[p]:off1[p + 1]], ids2[off2[p]:off2[p + 1]])
        u1, u2 = len(w[3]), len(w[4])
        assert r["lb"][p] == w[0] and r["l1"][p] == w[1] and r["l2"][p] == w[2]
        assert np.array_equal(r["argmin_rows"][off1[p]:off1[p] + u1], w[3])
        assert np.array_equal(r["argmin_cols"][off2[p]:off2[p] + u2], w[4])
        assert r["lb"][p] <= wmd[p] * (1 + 1e-5) + 1e-9
    e.close()


def test_device_entries_match_host_entry(eng_mod, dtab):
    import torch
    V = 3000
    table = workload.make_table(V, 300, seed=8)
    B = 70000                                                      # > one chunk: exercises both streams
    ids1, off1, ids2, off2 = workload.make_pairs(B, "yelp", "independent", V=V, seed=21)
    e = eng_mod.WMDEngine(table, distance_table=dtab)
    host, hst = e.wmd_pairs(ids1, off1, ids2, off2)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(a).to(dev)
    out, st = e.wmd_pairs_cuda(t(ids1), t(off1), t(ids2), t(off2), 20, 20)
    torch.cuda.synchronize()
    assert np.array_equal(out.cpu().numpy(), host) and np.array_equal(st.cpu().numpy(), hst)
    # padded entry: pad id 0 skipped, so shift rows by one through a token map
    e.set_token_map(np.concatenate([[-1], np.arange(V)]).astype(np.int32))
    A = np.zeros((B, 20), np.int64); Bm = np.zeros((B, 20), np.int64)
    for p in range(0, B, 97):                                      # a strided subset keeps the python loop short
        a = ids1[off1[p]:off1[p + 1]] + 1; b = ids2[off2[p]:off2[p + 1]] + 1
        A[p, :len(a)] = a; Bm[p, 20 - len(b):] = b                 # pads on either end
    sel = np.arange(0, B, 97)
    po, ps = e.wmd_pairs_padded(t(A[sel]), t(Bm[sel]), pad_id=0)
    torch.cuda.synchronize()
    assert np.array_equal(po.cpu().numpy(), host[sel]) and np.array_equal(ps.cpu().numpy(), hst[sel])
    e.close()


def test_symmetry_and_permutation_invariance_full_size(eng_mod):
    """Size-independent properties at the BASELINE shape (no oracle needed)."""
    table = workload.make_table(10000, 300, seed=0)
    B = 200000
    ids1, off1, ids2, off2 = workload.make_pairs(B, "yelp", "independent", V=10000, seed=1)
    e = eng_mod.WMDEngine(table)
    v12, s12 = e.wmd_pairs(ids1, off1, ids2, off2)
    v21, s21 = e.wmd_pairs(ids2, off2, ids1, off1)
    assert np.array_equal(s12, s21)
    fin = np.isfinite(v12)
    np.testing.assert_allclose(v12[fin], v21[fin], rtol=1e-12)
    # reversing the token order inside every document changes nothing (bags of words)
    lens = np.diff(off1)
    pos = np.arange(len(ids1)) - np.repeat(off1[:-1], lens)
    rev = np.repeat(off1[:-1], lens) + (np.repeat(lens, lens) - 1 - pos)
    vp, _ = e.wmd_pairs(ids1[rev], off1, ids2, off2)
    assert np.array_equal(vp, v12)
    # identical documents -> exactly 0.0 (or status 2 for one-token vocabularies)
    v0, s0 = e.wmd_pairs(ids1, off1, ids1, off1)
    assert np.all(v0 == 0.0) and set(np.unique(s0)) <= {0, 2}
    e.close()


@pytest.mark.parametrize("shape,B,d", [("yelp", 300, 300), ("book", 120, 100), ("fixed:40", 40, 64)])
def test_exact_mode_matches_lp(eng_mod, oracle, shape, B, d):
    """WMD_MODE_EXACT (additive): the un-quantised FP64 optimum against scipy HiGHS on the same nBOW
    histograms and float32 distances, 1e-9 relative; statuses as in the default mode; and within the
    1e-5 that the reference's 1e6 grid can move a value."""
    V = 800
    table = workload.make_table(V, d, seed=3)
    ids1, off1, ids2, off2 = workload.make_pairs(B, shape, "noised" if shape == "yelp" else "independent", V=V, seed=21)
    e = eng_mod.WMDEngine(table)
    got, st = e.wmd_pairs(ids1, off1, ids2, off2, mode="exact")
    ref, rst = e.wmd_pairs(ids1, off1, ids2, off2)                       # pyemd mode
    assert np.array_equal(st, rst)
    for p in range(B):
        want = oracle.wmd_exact_lp(table, ids1[off1[p]:off1[p + 1]], ids2[off2[p]:off2[p + 1]])
        if math.isinf(want) or want == 0.0:
            assert got[p] == want
        else:
            assert abs(got[p] - want) <= 1e-9 * max(1.0, want), (p, got[p], want)
            assert abs(got[p] - ref[p]) <= 2e-5 * max(1.0, want)
    e.close()


def test_async_and_device_forms_match_the_host_entries(eng_mod, oracle):
    """SURVEY 8(b) ABI: wmd_pairs_submit / wmd_pairs_wait, wmd_nbow_dev, wmd_rwmd_pairs_dev, wmd_workspace_bytes."""
    import torch
    V = 1500
    table = workload.make_table(V, 100, seed=4)
    ids1, off1, ids2, off2 = workload.make_pairs(3000, "book", "independent", V=V, seed=31)
    e = eng_mod.WMDEngine(table)
    want, wst = e.wmd_pairs(ids1, off1, ids2, off2)
    for _ in range(2):                                            # the second job reuses the pinned staging
        n = e.submit_pairs(ids1, off1, ids2, off2)
        with pytest.raises(RuntimeError, match="in flight"):
            e.wmd_pairs(ids1, off1, ids2, off2)
        got, st = e.wait_pairs(n)
        assert got.tobytes() == want.tobytes() and np.array_equal(st, wst)
    with pytest.raises(RuntimeError, match="no submitted job"):
        e.wait_pairs(1)
    # a token map on the device, rows passed through the flag
    tmap = np.concatenate([[-1, -1], np.arange(V)]).astype(np.int32)
    e.set_token_map(tmap)
    g2, s2 = e.wmd_pairs(ids1 + 2, off1, ids2 + 2, off2)
    g3, s3 = e.wmd_pairs(ids1, off1, ids2, off2, ids_are_rows=True)
    assert g2.tobytes() == want.tobytes() and g3.tobytes() == want.tobytes() and np.array_equal(s2, wst) and np.array_equal(s3, wst)
    e.set_token_map(None)
    dev = torch.device("cuda:0")
    t = lambda a: torch.from_numpy(a).to(dev)
    ml1, ml2 = int(np.diff(off1).max()), int(np.diff(off2).max())
    # nBOW on the device
    rows, counts, weights, uniq = e.nbow(ids1, off1)
    drows, dcounts, dweights, duniq = e.nbow_cuda(t(ids1), t(off1), ml1)
    torch.cuda.synchronize()
    assert np.array_equal(duniq.cpu().numpy(), uniq)
    for p in range(0, 3000, 7):
        a, u = off1[p], uniq[p]
        assert np.array_equal(drows.cpu().numpy()[a:a + u], rows[a:a + u]) and np.array_equal(dcounts.cpu().numpy()[a:a + u], counts[a:a + u])
        assert dweights.cpu().numpy()[a:a + u].tobytes() == weights[a:a + u].tobytes()
    # RWMD on the device (70 000 pairs: several chunks on both streams, argmins at absolute offsets)
    i1, o1, i2, o2 = workload.make_pairs(70000, "yelp", "independent", V=V, seed=33)
    h = e.rwmd_pairs(i1, o1, i2, o2)
    d = e.rwmd_pairs_cuda(t(i1), t(o1), t(i2), t(o2), 20, 20)
    torch.cuda.synchronize()
    for k in h:
        assert np.asarray(h[k]).tobytes() == d[k].cpu().numpy().tobytes(), k
    est, res = e.workspace_bytes(70000, 20, 20)
    assert res > table.nbytes and est >= V * V * 4 and est > 0
    est2, _ = e.workspace_bytes(70000, 256, 256)
    assert est2 > est
    e.close()


def test_distance_table_policy_follows_the_budget(eng_mod, monkeypatch):
    """Default policy: table on when V * V * 4 bytes fit the budget, direct path otherwise; same bits either way."""
    V = 1500
    table = workload.make_table(V, 64, seed=5)
    ids1, off1, ids2, off2 = workload.make_pairs(2000, "yelp", "noised", V=V, seed=41)
    e = eng_mod.WMDEngine(table)
    info = e.distance_table_info()
    assert info["enabled"] and not info["resident"] and info["bytes"] == V * V * 4     # built by the first scoring call
    a, sa = e.wmd_pairs(ids1, off1, ids2, off2)
    info = e.distance_table_info()
    assert info["resident"] and info["build_ms"] > 0
    e.close()
    monkeypatch.setenv("WMD_DTAB_BUDGET_MB", "1")                                       # 9 MB table, 1 MB budget
    e = eng_mod.WMDEngine(table)
    assert not e.distance_table_info()["enabled"]
    b, sb = e.wmd_pairs(ids1, off1, ids2, off2)
    assert not e.distance_table_info()["resident"]
    assert a.tobytes() == b.tobytes() and np.array_equal(sa, sb)
    e.close()


def test_fanout_duplicates_scores_and_status(eng_mod, dtab):
    """wmd_set_fanout on ONE GPU: every kernel family that stores a score or a status (fused kernel, list-mode nBOW,
    class A, wide classes; direct-path K1 / K3) must store the same value into the extra arrays -- the multi-GPU gather
    (sharding.PeerScores) rests on exactly this; two-rank coverage is in test_gpu_multirank.py."""
    import torch
    V = 900
    table = workload.make_table(V, 40, seed=8)
    table[11] = table[10]
    a1, o1, a2, o2 = workload.make_pairs(6000, "uniform:1-90", "independent", V=V, seed=4)        # short, long, wide classes
    docs1 = [[-1], [5], [10], [1, 2], []]; docs2 = [[1], [5], [11], [2, 1], [3]]                   # S1, S2, S4, zero distance, empty
    b1, p1 = workload.to_csr(docs1); b2, p2 = workload.to_csr(docs2)
    ids1 = np.concatenate([a1, b1]); off1 = np.concatenate([o1, o1[-1] + p1[1:]])
    ids2 = np.concatenate([a2, b2]); off2 = np.concatenate([o2, o2[-1] + p2[1:]])
    B = len(off1) - 1
    e = eng_mod.WMDEngine(table, distance_table=dtab)
    want, wst = e.wmd_pairs(ids1, off1, ids2, off2)
    dev = torch.device("cuda", e.device)
    # two extra copies, device entry
    x = [(torch.full((B,), -7.0, dtype=torch.float64, device=dev), torch.full((B,), -7, dtype=torch.int32, device=dev)) for _ in range(2)]
    e.set_fanout([t[0].data_ptr() for t in x], [t[1].data_ptr() for t in x])
    d = [torch.from_numpy(np.ascontiguousarray(v)).to(dev) for v in (ids1.astype(np.int32), off1, ids2.astype(np.int32), off2)]
    out, st = e.wmd_pairs_cuda(d[0], d[1], d[2], d[3], 90, 90)
    torch.cuda.synchronize()
    for o, s in x:
        assert torch.equal(o, out) and torch.equal(s, st)
    assert np.array_equal(out.cpu().numpy(), want) and np.array_equal(st.cpu().numpy(), wst)
    # host entry (chunked: the extra arrays are addressed per chunk), then off again
    for o, s in x:
        o.fill_(-7.0); s.fill_(-7)
    got, gst = e.wmd_pairs(ids1, off1, ids2, off2)
    torch.cuda.synchronize()
    for o, s in x:
        assert np.array_equal(o.cpu().numpy(), want) and np.array_equal(s.cpu().numpy(), wst)
    e.set_fanout()
    for o, s in x:
        o.fill_(-7.0); s.fill_(-7)
    e.wmd_pairs(ids1, off1, ids2, off2)
    torch.cuda.synchronize()
    assert float(x[0][0].max()) == -7.0 and int(x[0][1].max()) == -7
    with pytest.raises(RuntimeError):
        e.set_fanout([0] * 8, [0] * 8)                                     # more than WMD_MAX_FANOUT arrays
    # a peer buffer of this handle: alloc, use as a fan-out target, free
    ptr, handle = e.peer_alloc(12 * B)
    assert ptr != 0 and len(handle) == 64
    e.set_fanout([ptr], [ptr + 8 * B])
    e.wmd_pairs(ids1, off1, ids2, off2)
    e.set_fanout()
    torch.cuda.synchronize()

    class _Mem:
        def __init__(self, p, shape, typestr):
            self.__cuda_array_interface__ = {"data": (p, False), "shape": shape, "typestr": typestr, "version": 2}
    po = torch.as_tensor(_Mem(ptr, (B,), "<f8"), device=dev).clone(); ps = torch.as_tensor(_Mem(ptr + 8 * B, (B,), "<i4"), device=dev).clone()
    assert np.array_equal(po.cpu().numpy(), want) and np.array_equal(ps.cpu().numpy(), wst)
    e.peer_close(ptr, False)
    e.close()
