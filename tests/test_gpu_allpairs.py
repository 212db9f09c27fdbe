"""GPU parity of the all-pairs top-k mode (BASELINE.json configs[3]) against brute force.

The engine prunes with the relaxed lower bound and solves only the survivors; the oracle solves
every pair.  Indices must match exactly (ties broken by the lower document index) and distances
bit for bit, +inf rows included.
"""
import numpy as np
import pytest

from consistent__style_transfer_b200 import workload

pytestmark = pytest.mark.gpu


def _check(table, A, B, k, oracle, eng, row_begin=0, row_end=None):
    idsA, offA = A
    idsB, offB = B
    idx, dist, info = eng.allpairs_topk(idsA, offA, idsB, offB, k, row_begin, row_end)
    widx, wdist = oracle.allpairs_topk_bruteforce(table, idsA, offA, idsB, offB, k)
    hi = len(offA) - 1 if row_end is None else row_end
    widx, wdist = widx[row_begin:hi], wdist[row_begin:hi]
    assert dist.tobytes() == wdist.tobytes(), np.argwhere(dist != wdist)[:5]
    assert np.array_equal(idx, widx), np.argwhere(idx != widx)[:5]
    return info


@pytest.mark.parametrize("d", [300, 100])
def test_allpairs_topk_matches_bruteforce(oracle, d):
    from consistent__style_transfer_b200.engine import WMDEngine
    V = 600
    table = workload.make_table(V, d, seed=2)
    idsA, offA, idsB, offB = workload.make_pairs(150, "yelp", "independent", V=V, seed=11)
    idsB, offB = idsB[:offB[140]], offB[:141]
    eng = WMDEngine(table)
    info = _check(table, (idsA, offA), (idsB, offB), 8, oracle, eng)
    # the bound must actually prune: far fewer exact solves than pairs
    assert info["bounds"] == 150 * 140
    assert info["exact_round1"] + info["exact_round2"] < 0.5 * info["bounds"]
    # a row block (what one rank of a multi-GPU job computes)
    _check(table, (idsA, offA), (idsB, offB), 5, oracle, eng, row_begin=37, row_end=101)
    eng.close()


def test_allpairs_wide_corpus_fast_selection(oracle):
    """|B| >= 256 takes the two-pass k-th selection (thread minima + ranked short list); heavy ties in the
    bound (many identical documents) take the radix fallback."""
    from consistent__style_transfer_b200.engine import WMDEngine
    V = 500
    table = workload.make_table(V, 32, seed=4)
    idsA, offA, idsB, offB = workload.make_pairs(330, "yelp", "independent", V=V, seed=17)
    idsA, offA = idsA[:offA[60]], offA[:61]
    eng = WMDEngine(table)
    _check(table, (idsA, offA), (idsB, offB), 7, oracle, eng)
    # 2100 copies of three documents: thousands of bounds tie at the k-th value
    docs = [[1, 2, 3], [4, 5], [6]] * 700 + [[7, 8, 9, 10]] * 3
    ids, off = workload.to_csr(docs)
    qa, qo = workload.to_csr([[1, 2, 3], [4, 6], [7, 8, 9, 10], [11]])
    _check(table, (qa, qo), (ids, off), 5, oracle, eng)
    eng.close()


def test_allpairs_self_join_duplicates_oov_and_empty_docs(oracle):
    """A == B with duplicated documents (zero distances, index tie-breaks), out-of-vocabulary ids and
    documents that are empty after OOV removal (+inf against everything)."""
    from consistent__style_transfer_b200.engine import WMDEngine
    V = 300
    table = workload.make_table(V, 64, seed=5)
    rng = np.random.default_rng(7)
    docs = []
    for _ in range(90):
        n = int(rng.integers(1, 13))
        docs.append(list(rng.integers(0, V, size=n)))
    docs[10] = list(docs[3]); docs[11] = list(docs[3])[::-1]       # same bag, different order
    docs[20] = [-1, -1]; docs[21] = []                              # empty after the OOV filter / raw empty
    docs[30] = docs[30] + [-1, -1]                                  # OOV tokens (row -1) are dropped
    docs[40] = [7]; docs[41] = [7]; docs[42] = [7, 7, 7]            # one-token unions -> 0.0 (status 2)
    ids, off = workload.to_csr(docs)
    eng = WMDEngine(table)
    _check(table, (ids, off), (ids, off), 6, oracle, eng)
    eng.close()


def test_allpairs_k_equals_corpus_and_book_shape(oracle):
    from consistent__style_transfer_b200.engine import WMDEngine
    V = 400
    table = workload.make_table(V, 100, seed=9)
    idsA, offA, idsB, offB = workload.make_pairs(40, "book", "independent", V=V, seed=13)
    idsB, offB = idsB[:offB[24]], offB[:25]
    eng = WMDEngine(table)
    _check(table, (idsA, offA), (idsB, offB), 24, oracle, eng)      # k == |B|: a full sort of every row
    eng.close()


def test_allpairs_unnormalised_table_long_documents_and_device_output(oracle):
    """An un-normalised embedding table (row norms ~10: distances of 10-20, where an absolute pruning margin
    sized for unit vectors would be too small) with documents of up to 120 tokens; and the device-output entry
    (wmd_allpairs_topk_dev) returns the same bits as the host entry."""
    import torch
    from consistent__style_transfer_b200.engine import WMDEngine
    V = 700
    rng = np.random.default_rng(19)
    table = (rng.standard_normal((V, 48)) * 1.5).astype(np.float32)
    idsA, offA, idsB, offB = workload.make_pairs(260, "uniform:1-120", "independent", V=V, seed=29)
    idsA, offA = idsA[:offA[40]], offA[:41]
    eng = WMDEngine(table)
    _check(table, (idsA, offA), (idsB, offB), 9, oracle, eng)
    idx, dist, _ = eng.allpairs_topk(idsA, offA, idsB, offB, 9, 5, 33)
    didx, ddist, info = eng.allpairs_topk_cuda(idsA, offA, idsB, offB, 9, 5, 33)
    torch.cuda.synchronize()
    assert didx.is_cuda and np.array_equal(didx.cpu().numpy(), idx) and ddist.cpu().numpy().tobytes() == dist.tobytes()
    assert info["bounds"] == 28 * 260
    eng.close()
