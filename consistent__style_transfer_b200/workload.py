"""Deterministic synthetic inputs of the shapes BASELINE.json names.

Nothing here computes WMD: this only draws embedding tables and token-id documents so that
tests, ``bench.py`` and the CPU baseline all see the same bytes.

* table: ``default_rng(0).standard_normal((V, d), float32)``, rows L2-normalised in float32 the
  ``init_sims(replace=True)`` way (/root/reference/src/wmd.py:54).
* documents: token ranks from a Zipf-Mandelbrot law ``p(r) ~ 1/(r + 2.7)`` over the V rows.
* ``yelp``  lengths ``1 + binomial(19, 0.45)`` (1..20, mean 9.55; real Yelp mean 9.2)
* ``book``  lengths ``clip(1 + binomial(63, 0.22), 1, 64)`` (mean ~14.9; real book 14.6)
* ``fixed:L`` both sides exactly L tokens (length sweep, BASELINE config 5)
* ``uniform:A-B`` lengths uniform in [A, B], drawn per side (mixed solver classes and cost-stage kinds in one launch)
* variant ``independent`` (doc2 drawn independently: worst case, little cancellation) or
  ``noised`` (doc2 = doc1 with each token moved w.p. 0.15 to a random other document of its
  batch of 256, after /root/reference/src/data_util.py:32-54).
"""
from __future__ import annotations

import numpy as np


def make_table(V: int = 10_000, d: int = 300, seed: int = 0) -> np.ndarray:
    rng = np.random.default_rng(seed)
    E = rng.standard_normal((V, d), dtype=np.float32)
    nrm = np.sqrt((E ** 2).sum(-1, dtype=np.float32)).astype(np.float32)
    E /= nrm[:, None]
    return np.ascontiguousarray(E, dtype=np.float32)


def _zipf_cdf(V: int) -> np.ndarray:
    p = 1.0 / (np.arange(V, dtype=np.float64) + 2.7)
    p /= p.sum()
    return np.cumsum(p)


def _lengths(rng, shape: str, B: int) -> np.ndarray:
    if shape == "yelp":
        return 1 + rng.binomial(19, 0.45, size=B)
    if shape == "book":
        return np.clip(1 + rng.binomial(63, 0.22, size=B), 1, 64)
    if shape.startswith("fixed:"):
        return np.full(B, int(shape.split(":")[1]), dtype=np.int64)
    if shape.startswith("uniform:"):
        lo, hi = (int(x) for x in shape.split(":")[1].split("-"))
        return rng.integers(lo, hi + 1, size=B)
    raise ValueError(shape)


def _draw_docs(rng, cdf, lens):
    total = int(lens.sum())
    ids = np.searchsorted(cdf, rng.random(total), side="right").astype(np.int32)
    np.minimum(ids, len(cdf) - 1, out=ids)
    off = np.zeros(len(lens) + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    return ids, off


def make_pairs(B: int, shape: str = "yelp", variant: str = "independent", V: int = 10_000,
               seed: int = 1, batch: int = 256):
    """Returns CSR (ids1 int32, off1 int64, ids2 int32, off2 int64) of B document pairs."""
    rng = np.random.default_rng(seed)
    cdf = _zipf_cdf(V)
    ids1, off1 = _draw_docs(rng, cdf, _lengths(rng, shape, B))
    if variant == "independent":
        ids2, off2 = _draw_docs(rng, cdf, _lengths(rng, shape, B))
        return ids1, off1, ids2, off2
    if variant != "noised":
        raise ValueError(variant)
    # transfer_noise: each token leaves its sentence w.p. 0.15 and lands in a sentence of the
    # same batch drawn with probability proportional to sentence length, at a random position.
    docs2 = []
    for b0 in range(0, B, batch):
        b1 = min(B, b0 + batch)
        sents = [list(ids1[off1[p]:off1[p + 1]]) for p in range(b0, b1)]
        lens = np.array([len(s) for s in sents], dtype=np.float64)
        bag, kept = [], []
        for s in sents:
            mv = rng.random(len(s)) < 0.15
            kept.append([t for t, m in zip(s, mv) if not m])
            bag.extend(t for t, m in zip(s, mv) if m)
        dest = rng.choice(len(sents), size=len(bag), p=lens / lens.sum())
        for t, k in zip(bag, dest):
            pos = int(rng.integers(0, len(kept[k]) + 1))
            kept[k].insert(pos, t)
        docs2.extend(kept)
    lens2 = np.array([len(s) for s in docs2], dtype=np.int64)
    off2 = np.zeros(B + 1, np.int64)
    np.cumsum(lens2, out=off2[1:])
    ids2 = np.fromiter((t for s in docs2 for t in s), dtype=np.int32, count=int(lens2.sum()))
    return ids1, off1, ids2, off2


def to_csr(docs):
    """list of id lists -> (ids int32, off int64)"""
    lens = np.fromiter((len(d) for d in docs), dtype=np.int64, count=len(docs))
    off = np.zeros(len(docs) + 1, np.int64)
    np.cumsum(lens, out=off[1:])
    ids = np.fromiter((t for d in docs for t in d), dtype=np.int32, count=int(off[-1]))
    return ids, off
