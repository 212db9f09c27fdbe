"""Pins oracle/emd_hat.c (restatement of pyemd 0.5.1 emd_hat_gd_metric<double>, spec S6 of
SURVEY.md 8(c)) with pyemd's own known-answer vectors and two independent exact solvers."""
import numpy as np
import pytest


# ---- pyemd known-answer vectors (upstream test_pyemd.py) ----------
def test_kat_unequal_mass(oracle):
    assert oracle.emd([0.0, 1.0], [5.0, 3.0], [[0.0, 0.5], [0.5, 0.0]]) == pytest.approx(3.5, abs=1e-5)


def test_kat_identical(oracle):
    assert oracle.emd([1.0, 1.0], [1.0, 1.0], [[0.0, 1.0], [1.0, 0.0]]) == 0.0


def test_kat_extra_mass_penalty(oracle):
    D = [[0.0, 1.0, 1.0, 2.0], [1.0, 0.0, 2.0, 1.0], [1.0, 2.0, 0.0, 1.0], [2.0, 1.0, 1.0, 0.0]]
    v = oracle.emd([0.0, 2.0, 1.0, 2.0], [2.0, 1.0, 2.0, 1.0], D, extra_mass_penalty=2.5)
    assert v == pytest.approx(4.5, abs=1e-5)      # upstream asserts to 5 decimals (the 1e-6 grid)


def test_kat_sti_two_bins(oracle):
    # /root/reference/evaluate/auto/transfer_intensity.py:8-11: emd(p, q, ones((2,2)))
    v = oracle.emd([0.9, 0.1], [0.2, 0.8], np.ones((2, 2)))
    assert v == pytest.approx(0.7, abs=1e-6)


# ---- independent solver 1: networkx network simplex on the quantised integers ------------
def _nx_integer_opt(iP, iQ, iC):
    import networkx as nx
    sP, sQ = int(iP.sum()), int(iQ.sum())
    if sQ > sP:
        iP, iQ, iC = iQ, iP, iC.T
        sP, sQ = sQ, sP
    G = nx.DiGraph()
    src = [i for i in range(len(iP)) if iP[i] > 0]
    dst = [j for j in range(len(iQ)) if iQ[j] > 0]
    if not dst:
        return 0
    for i in src:
        G.add_node(("s", i), demand=-int(iP[i]))
    for j in dst:
        G.add_node(("t", j), demand=int(iQ[j]))
    G.add_node("dump", demand=sP - sQ)          # surplus leaves at zero cost
    for i in src:
        G.add_edge(("s", i), "dump", weight=0)
        for j in dst:
            G.add_edge(("s", i), ("t", j), weight=int(iC[i, j]))
    cost, _ = nx.network_simplex(G)
    return int(cost)


def _random_problem(rng, n, density=0.7):
    x = rng.standard_normal((n, 8)).astype(np.float32)
    D = np.sqrt(((x[:, None, :] - x[None, :, :]) ** 2).sum(-1)).astype(np.float64)
    D = (D + D.T) / 2
    np.fill_diagonal(D, 0.0)
    c1 = rng.integers(0, 4, n) * (rng.random(n) < density)
    c2 = rng.integers(0, 4, n) * (rng.random(n) < density)
    if c1.sum() == 0:
        c1[0] = 1
    if c2.sum() == 0:
        c2[-1] = 2
    return c1 / float(c1.sum()), c2 / float(c2.sum()), D


@pytest.mark.parametrize("n", [2, 3, 5, 9, 17, 33, 96, 200])
def test_integer_stage_vs_network_simplex(oracle, n):
    # up to the problem sizes of the GPU's wide solver classes (long documents): the oracle those are checked against
    # is itself checked against an independent exact solver there
    rng = np.random.default_rng(100 + n)
    for _ in range(25 if n < 20 else 6 if n < 64 else 2):
        d1, d2, D = _random_problem(rng, n)
        iP, iQ, iC = oracle.emd_quantise(d1, d2, D)
        assert oracle.emd_integral(iP, iQ, iC, 0) == _nx_integer_opt(iP, iQ, iC)


# ---- independent solver 2: HiGHS LP on the un-quantised problem --------------------------
def _lp_exact(d1, d2, D):
    from scipy.optimize import linprog
    n = len(d1)
    P, Q = d1.copy(), d2.copy()
    c = D.reshape(-1)
    A_eq, b_eq = [], []
    for i in range(n):
        row = np.zeros((n, n)); row[i, :] = 1; A_eq.append(row.reshape(-1)); b_eq.append(P[i])
    for j in range(n):
        row = np.zeros((n, n)); row[:, j] = 1; A_eq.append(row.reshape(-1)); b_eq.append(Q[j])
    res = linprog(c, A_eq=np.array(A_eq), b_eq=np.array(b_eq), bounds=(0, None), method="highs")
    assert res.status == 0
    return res.fun


@pytest.mark.parametrize("n", [3, 6, 12, 20])
def test_quantised_close_to_exact_lp(oracle, n):
    # pyemd's 1e-6 grid moves the value by ~1e-6 relative (max ~1e-5)
    rng = np.random.default_rng(7 + n)
    for _ in range(10):
        d1, d2, D = _random_problem(rng, n)
        exact = _lp_exact(d1, d2, D)
        got = oracle.emd(d1, d2, D)
        assert got == pytest.approx(exact, rel=5e-5, abs=1e-9)


def test_symmetry_and_zero(oracle):
    rng = np.random.default_rng(3)
    for _ in range(20):
        d1, d2, D = _random_problem(rng, 7)
        assert oracle.emd(d1, d2, D) == pytest.approx(oracle.emd(d2, d1, D), rel=1e-12)
        assert oracle.emd(d1, d1, D) == 0.0


def test_one_by_n_closed_form(oracle):
    rng = np.random.default_rng(5)
    n = 6
    _, d2, D = _random_problem(rng, n)
    d1 = np.zeros(n); d1[0] = 1.0
    d2[0] = 0.0; d2 /= d2.sum()
    assert oracle.emd(d1, d2, D) == pytest.approx(float((d2 * D[0]).sum()), rel=3e-6)
