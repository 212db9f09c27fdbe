"""The algorithmic choices of the GPU transport solver (solve.cuh), checked on the CPU: the python model of its search
(tools/solver_model.py) must reach the oracle's exact integer optimum whichever side supplies the rows and whether or not
a search continues after an augmentation that leaves its tree intact.  The GPU parity tests check the kernels themselves;
this pins the method."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("solver_model", os.path.join(ROOT, "tools", "solver_model.py"))
solver_model = importlib.util.module_from_spec(spec)
spec.loader.exec_module(solver_model)


def _balanced(iP, iQ, iC):
    """Rows = the heavier side's residual nodes, columns = the lighter side's plus the surplus as a zero-cost dummy column
    (what emd_solve_small_kernel sets up before it decides which side becomes the rows)."""
    iP, iQ, iC = np.asarray(iP, np.int64), np.asarray(iQ, np.int64), np.asarray(iC, np.int64)
    if iP.sum() < iQ.sum():
        iP, iQ, iC = iQ, iP, iC.T
    r, c = np.nonzero(iP > 0)[0], np.nonzero(iQ > 0)[0]
    if len(r) == 0 or len(c) == 0:
        return None
    C, s, t = iC[np.ix_(r, c)], iP[r], iQ[c]
    diff = int(s.sum() - t.sum())
    if diff > 0:
        C = np.hstack([C, np.zeros((len(r), 1), np.int64)]); t = np.append(t, diff)
    return C, s, t


def _random_problem(rng, n, maxcount, dim):
    x = rng.standard_normal((n, dim)).astype(np.float32)
    if dim <= 2:
        x = np.round(x)                                  # lattice points: many tied costs
    D = np.sqrt(((x[:, None, :] - x[None, :, :]) ** 2).sum(-1)).astype(np.float64)
    np.fill_diagonal(D, 0.0)
    c1 = rng.integers(0, maxcount + 1, n); c2 = rng.integers(0, maxcount + 1, n)
    if c1.sum() == 0:
        c1[0] = 1
    if c2.sum() == 0:
        c2[-1] = 2
    return c1 / float(c1.sum()), c2 / float(c2.sum()), D


@pytest.mark.parametrize("n,maxcount,dim", [(2, 1, 8), (3, 3, 8), (6, 2, 1), (9, 3, 8), (14, 1, 2), (20, 4, 8), (33, 2, 8)])
def test_search_variants_reach_the_exact_optimum(oracle, n, maxcount, dim):
    rng = np.random.default_rng(1000 * n + maxcount)
    checked = 0
    for _ in range(20 if n < 30 else 5):
        d1, d2, D = _random_problem(rng, n, maxcount, dim)
        if D.max() == 0:
            continue
        iP, iQ, iC = oracle.emd_quantise(d1, d2, D)
        want = oracle.emd_integral(iP, iQ, iC, 0)
        prob = _balanced(iP, iQ, iC)
        if prob is None:
            assert want == 0
            continue
        C, s, t = prob
        flipped = (C.T.copy(), t.copy(), s.copy())
        for name, p, cont, red in (("rows = supplying side", (C, s, t), False, False), ("rows = other side", flipped, False, False),
                                   ("supplying side, searches continue", (C, s, t), True, False),
                                   ("other side, searches continue", flipped, True, False),
                                   ("supplying side, reduced-cost start + greedy pass", (C, s, t), True, True),
                                   ("other side, reduced-cost start + greedy pass", flipped, True, True)):
            got, counts = solver_model.solve(*p, cont, red)
            assert got == want, (name, n, got, want)
        checked += 1
    assert checked > 0
