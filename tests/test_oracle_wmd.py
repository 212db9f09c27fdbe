"""Pins oracle/wmd_oracle.{py,c} (gensim wmdistance restatement, spec S1..S5 of SURVEY.md 8(c))."""
import math

import numpy as np
import pytest

from consistent__style_transfer_b200 import workload


@pytest.mark.parametrize("d", [1, 5, 7, 8, 9, 50, 100, 127, 128, 129, 144, 256, 257, 300, 304, 512, 1000])
def test_c_distance_is_numpy_bit_for_bit(oracle, d):
    # gensim: sqrt(np_sum((self[t1] - self[t2])**2)) on float32 rows -> numpy pairwise order
    rng = np.random.default_rng(d)
    for _ in range(300):
        a = rng.standard_normal(d).astype(np.float32)
        b = rng.standard_normal(d).astype(np.float32)
        want = np.sqrt(np.sum((a - b) ** 2))
        assert want.dtype == np.float32
        got = oracle.dist_f32(a, b)
        assert got.tobytes() == want.tobytes()


def _string_vocab(V, rng):
    words = set()
    while len(words) < V:
        n = int(rng.integers(1, 7))
        words.add("".join(chr(int(c)) for c in rng.integers(97, 123, n)))
    words = list(words)
    rng.shuffle(words)
    return words


@pytest.mark.parametrize("d,shape", [(100, "yelp"), (300, "yelp"), (100, "book")])
def test_python_loop_equals_c_batch(oracle, d, shape):
    rng = np.random.default_rng(11)
    V = 400
    words = _string_vocab(V, rng)
    table = workload.make_table(V, d, seed=2)
    kv = oracle.KeyedVectorsOracle(words, table)
    rank = oracle.string_rank(words)
    B = 60
    ids1, off1, ids2, off2 = workload.make_pairs(B, shape, "noised", V=V, seed=5, batch=20)
    # sprinkle OOV tokens
    ids1 = ids1.copy(); ids1[rng.random(len(ids1)) < 0.05] = -1
    got, st = oracle.batch_wmd(table, ids1, off1, ids2, off2, rank=rank)
    for p in range(B):
        d1 = [words[t] if t >= 0 else "<<oov>>" for t in ids1[off1[p]:off1[p + 1]]]
        d2 = [words[t] for t in ids2[off2[p]:off2[p + 1]]]
        want = kv.wmdistance(d1, d2)
        if math.isinf(want):
            assert math.isinf(got[p]) and st[p] in (1, 3)
        else:
            assert got[p] == want, (p, got[p], want)


def test_early_outs(oracle):
    table = workload.make_table(50, 16, seed=3)
    table[7] = table[6]                      # two different tokens with identical vectors
    words = ["w%02d" % i for i in range(50)]
    kv = oracle.KeyedVectorsOracle(words, table)
    assert kv.wmdistance(["zzz"], ["w01"]) == float("inf")             # S1: empty after OOV removal
    assert kv.wmdistance(["w01"], []) == float("inf")
    assert kv.wmdistance(["w01", "w01"], ["w01"]) == 0.0               # S2: one-token vocabulary
    assert kv.wmdistance(["w06"], ["w07"]) == float("inf")             # S4: all-zero matrix
    assert kv.wmdistance(["w01", "w02"], ["w02", "w01"]) == 0.0        # identical bags
    ids1, off1 = workload.to_csr([[-1], [1], [1, 1], [6], [1, 2]])
    ids2, off2 = workload.to_csr([[1], [], [1], [7], [2, 1]])
    v, st = oracle.batch_wmd(table, ids1, off1, ids2, off2)
    assert list(st) == [1, 1, 2, 3, 0]
    assert math.isinf(v[0]) and math.isinf(v[1]) and v[2] == 0.0 and math.isinf(v[3]) and v[4] == 0.0


def test_single_token_docs_give_vector_distance(oracle):
    table = workload.make_table(50, 100, seed=4)
    ids1, off1 = workload.to_csr([[3]]); ids2, off2 = workload.to_csr([[9]])
    v, _ = oracle.batch_wmd(table, ids1, off1, ids2, off2)
    want = float(np.sqrt(np.sum((table[3] - table[9]) ** 2)))
    assert v[0] == pytest.approx(want, rel=2e-6)       # one 1e6-grid rounding of mass and cost


def test_wrapper_fallbacks(oracle):
    # /root/reference/src/wmd.py:37-44
    table = workload.make_table(20, 8, seed=5)
    words = ["t%d" % i for i in range(20)]

    class Tok:
        def ids_to_tokens(self, ids):
            return [words[i] if 0 <= i < 20 else None for i in ids]

    w = oracle.WMDdistanceOracle(oracle.KeyedVectorsOracle(words, table))
    lab = w.cal_wmd_label([[], [1, 2, 3], [99, 98], [4, 5]], [[1, 2], [], [1], [5, 4]], Tok())
    assert lab[0] == 2.0 and lab[1] == 3.0          # raw empty list -> max(len)
    assert lab[2] == 1.5                            # inf (all OOV) -> (len1+len2)/2
    assert lab[3] == 0.0


def test_symmetry_permutation_and_bounds(oracle):
    rng = np.random.default_rng(9)
    table = workload.make_table(300, 100, seed=6)
    ids1, off1, ids2, off2 = workload.make_pairs(200, "yelp", "independent", V=300, seed=8)
    v12, _ = oracle.batch_wmd(table, ids1, off1, ids2, off2)
    v21, _ = oracle.batch_wmd(table, ids2, off2, ids1, off1)
    np.testing.assert_allclose(v12, v21, rtol=1e-12)
    # permutation invariance: shuffle tokens inside each doc
    ids1p = ids1.copy()
    for p in range(200):
        seg = ids1p[off1[p]:off1[p + 1]]
        rng.shuffle(seg)
    vp, _ = oracle.batch_wmd(table, ids1p, off1, ids2, off2)
    assert np.array_equal(vp, v12)
    # RWMD <= WMD (up to the 1e-6 grid)
    for p in range(0, 200, 7):
        r = oracle.rwmd_pair(table, ids1[off1[p]:off1[p + 1]], ids2[off2[p]:off2[p + 1]])
        assert r[0] <= v12[p] * (1 + 1e-5) + 1e-9


def test_exact_lp_oracle_close_to_quantised(oracle):
    """The additive exact mode's checker (HiGHS LP) against the quantised reference value: they differ
    only by what the 1e6 grid can move (median 1e-6, max ~1e-5 relative)."""
    import numpy as np
    from consistent__style_transfer_b200 import workload
    V = 300
    table = workload.make_table(V, 50, seed=8)
    ids1, off1, ids2, off2 = workload.make_pairs(60, "yelp", "independent", V=V, seed=2)
    q, st = oracle.batch_wmd(table, ids1, off1, ids2, off2)
    worst = 0.0
    for p in range(60):
        lp = oracle.wmd_exact_lp(table, ids1[off1[p]:off1[p + 1]], ids2[off2[p]:off2[p + 1]])
        if np.isfinite(q[p]) and q[p] > 0:
            worst = max(worst, abs(lp - q[p]) / q[p])
        else:
            assert lp == q[p]
    assert 0.0 < worst < 3e-5


def test_workload_shapes_are_deterministic_and_in_range():
    # bench.py, the tests and the CPU baseline must all see the same bytes for a (shape, variant, seed)
    for shape, lo, hi in (("yelp", 1, 20), ("book", 1, 64), ("fixed:37", 37, 37), ("uniform:3-90", 3, 90)):
        a = workload.make_pairs(300, shape, "independent", V=500, seed=4)
        b = workload.make_pairs(300, shape, "independent", V=500, seed=4)
        for x, y in zip(a, b):
            assert x.dtype == y.dtype and np.array_equal(x, y)
        for off in (a[1], a[3]):
            lens = np.diff(off)
            assert off[0] == 0 and lens.min() >= lo and lens.max() <= hi
        assert a[0].dtype == np.int32 and a[1].dtype == np.int64
        assert a[0].min() >= 0 and a[0].max() < 500
    n1 = workload.make_pairs(256, "yelp", "noised", V=500, seed=4)
    assert np.diff(n1[3]).sum() == np.diff(n1[1]).sum()          # transfer_noise moves tokens, it never drops them
