"""GPU parity of the batched pyemd.emd entry (wmd_emd_batch_host) against the CPU oracle
(oracle/emd_hat.c:emd_hat_gd_metric_double) and pyemd's own known-answer vectors, plus the
transfer-intensity drop-in (reference evaluate/auto/transfer_intensity.py)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def eng():
    from consistent__style_transfer_b200.engine import WMDEngine
    e = WMDEngine(np.ones((1, 1), np.float32))
    yield e
    e.close()


def test_pyemd_known_answers(eng):
    kats = json.load(open(os.path.join(HERE, "golden", "pyemd_known_answers.json")))
    for c in kats:
        P = np.array([c["p"]], np.float64); Q = np.array([c["q"]], np.float64)
        D = np.array(c["D"], np.float64)
        got = eng.emd_batch(P, Q, D, c["penalty"])[0]
        assert round(abs(got - c["want"]), c["decimals"]) == 0, (c, got)     # assertAlmostEqual(places=decimals), as upstream


@pytest.mark.parametrize("n", [1, 2, 3, 7, 16, 31])
def test_random_problems_bit_exact(eng, oracle, n):
    rng = np.random.default_rng(100 + n)
    B = 400
    P = rng.random((B, n)); Q = rng.random((B, n))
    P[rng.random((B, n)) < 0.3] = 0.0; Q[rng.random((B, n)) < 0.3] = 0.0
    P[:, 0] += 1e-3; Q[:, -1] += 1e-3                               # positive total mass (pyemd's precondition)
    half = B // 2
    P[:half] /= P[:half].sum(1, keepdims=True); Q[:half] /= Q[:half].sum(1, keepdims=True)   # equal masses: nBOW-like
    X = rng.standard_normal((B, n, 5))
    D = np.sqrt(((X[:, :, None, :] - X[:, None, :, :]) ** 2).sum(-1)) + 0.01                  # symmetric, positive max
    for emp in (-1.0, 0.0, 2.5):
        got = eng.emd_batch(P, Q, D, emp)
        want = np.array([oracle.emd(P[b], Q[b], D[b], emp) for b in range(B)])
        assert got.tobytes() == want.tobytes(), (n, emp, np.nonzero(got != want)[0][:5])
    # one matrix shared by the whole batch
    got = eng.emd_batch(P, Q, D[0])
    want = np.array([oracle.emd(P[b], Q[b], D[0]) for b in range(B)])
    assert got.tobytes() == want.tobytes()


def test_asymmetric_matrix_follows_upstream_indexing(eng, oracle):
    """pyemd reads C[supplier bin][consumer bin] even after the histograms swap roles."""
    rng = np.random.default_rng(5)
    B, n = 200, 6
    P = rng.random((B, n)); Q = rng.random((B, n)) * rng.choice([0.5, 2.0], size=(B, 1))
    D = rng.random((B, n, n)) + 0.05
    got = eng.emd_batch(P, Q, D)
    want = np.array([oracle.emd(P[b], Q[b], D[b]) for b in range(B)])
    assert got.tobytes() == want.tobytes()


def test_transfer_intensity_dropin(eng, oracle):
    from consistent__style_transfer_b200 import transfer_intensity as ti

    class FakeFasttext:                                              # predict(sequence, k) -> (labels, probabilities), unsorted
        labels = ["__label__0", "__label__1"]

        def predict(self, sequence, k):
            h = (hash(sequence) % 1000) / 1000.0
            return ("__label__1", "__label__0"), np.array([h, 1.0 - h])

    assert ti.calculate_emd([0.9, 0.1], [0.2, 0.8], eng) == oracle.emd([0.9, 0.1], [0.2, 0.8], np.ones((2, 2)))
    assert abs(ti.calculate_emd([0.9, 0.1], [0.2, 0.8], eng) - 0.7) < 1e-6
    assert ti.calculate_direction_corrected_emd([0.9, 0.1], [0.2, 0.8], 0, eng) < 0 < \
        ti.calculate_direction_corrected_emd([0.9, 0.1], [0.2, 0.8], 1, eng)
    ins = ["and the cleaning is way over priced .", "i hate the cornbread appetizer .", "ok"]
    outs = ["and the cleaning is way perfectly priced .", "i love the cornbread appetizer .", "ok"]
    got = ti.calculate_STIs(ins, outs, [1, 0, 1], FakeFasttext(), eng)
    m = FakeFasttext()
    for s_in, s_out, tgt, g in zip(ins, outs, [1, 0, 1], got):
        pi = np.array([1.0 - (hash(s_in) % 1000) / 1000.0, (hash(s_in) % 1000) / 1000.0])
        po = np.array([1.0 - (hash(s_out) % 1000) / 1000.0, (hash(s_out) % 1000) / 1000.0])
        want = oracle.emd(pi, po, np.ones((2, 2))) * (1 if po[tgt] >= pi[tgt] else -1)
        assert g == want
