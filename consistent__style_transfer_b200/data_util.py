"""Sentence noising of the pretrain label pipeline on flat (CSR) arrays.

Stands where `align`, `transfer_noise` and `rand_perm` of the reference's src/data_util.py:25-74 stand (same
names, arguments and return shapes, so `collate_pretrain` and user code keep working), but is written for the
batch, not the token: a batch is flattened once into one id array plus offsets, every Bernoulli draw of the batch
comes from ONE `np.random.uniform` call, and kept / moved tokens are split with boolean masks.

Compatibility contract (checked by tests/test_data_util_cpu.py against outputs of the reference's own source,
tests/golden/noise_cases.json.gz): with the same `np.random.seed` / `random.seed`, the results are identical to
the reference's.  That pins the ORDER in which the two global generators are consumed --

* `transfer_noise`: `len(s)` uniforms per sentence in batch order (numpy's legacy generator hands out the same
  doubles whether they are asked for sentence by sentence or all at once), then one `np.random.choice` over the
  sentences with probability proportional to the ORIGINAL lengths, then one `random.randint(0, current length of
  the target)` per moved token, in bag order;
* `rand_perm`: one uniform per token of the batch, then one `random.shuffle` of the selected tokens.

The reference's `np.float` (data_util.py:44, removed in numpy 1.24) meant the builtin float: float64 here.
`transfer_noise_cuda` / `rand_perm_cuda` are the device-resident counterparts for padded CUDA batches (new; they
draw from a torch generator, so they match the reference in distribution, not in stream).
"""
from __future__ import annotations

import random
from typing import List, Sequence, Tuple

import numpy as np


def flatten(sentences: Sequence[Sequence[int]]) -> Tuple[np.ndarray, np.ndarray]:
    """list of id lists -> (flat int64 ids, int64 offsets[len + 1]); both writable"""
    from .engine import docs_to_csr                      # the C packer when it is built, numpy otherwise
    flat, off = docs_to_csr(sentences, np.int64)
    return (flat if flat.flags.writeable else flat.copy()), (off if off.flags.writeable else off.copy())


def split(flat: np.ndarray, off: np.ndarray) -> List[list]:
    """inverse of `flatten`: python lists of python ints"""
    vals = flat.tolist()
    bounds = off.tolist()
    return [vals[bounds[i]:bounds[i + 1]] for i in range(len(bounds) - 1)]


def align(sentences, pad_value, max_len=None):
    """Pads (and truncates) every sentence to `max_len` (default: the longest).  Returns
    (padded sentences, truncated lengths, max_len) like data_util.py:25-30."""
    lens = [len(s) for s in sentences]
    if max_len is None:
        max_len = max(lens)
    lengths = [l if l < max_len else max_len for l in lens]
    padded = [list(s[:max_len]) + [pad_value] * (max_len - l) for s, l in zip(sentences, lengths)]
    return padded, lengths, max_len


def align_array(sentences, pad_value: int, max_len=None) -> np.ndarray:
    """`align` straight into an int64 [B, max_len] array (what the collate function turns into a tensor)."""
    flat, off = flatten(sentences)
    lens = np.diff(off)
    width = int(lens.max()) if max_len is None else int(max_len)
    grid = np.full((len(sentences), width), pad_value, np.int64)
    col = np.arange(flat.shape[0]) - np.repeat(off[:-1], lens)
    keep = col < width
    grid[np.repeat(np.arange(len(sentences)), lens)[keep], col[keep]] = flat[keep]
    return grid


def transfer_noise(sentences, p):
    """Every token leaves its sentence with probability `p` and lands at a random position of a sentence of the
    batch drawn with probability proportional to the original sentence lengths (data_util.py:32-54)."""
    flat, off = flatten(sentences)
    n = len(sentences)
    lens = np.diff(off)
    leaves = np.random.uniform(size=flat.shape[0]) < p                     # the batch's Bernoulli draws, in token order
    owner = np.repeat(np.arange(n), lens)
    stay_off = np.zeros(n + 1, np.int64)
    np.cumsum(np.bincount(owner[~leaves], minlength=n), out=stay_off[1:])
    noised = split(flat[~leaves], stay_off)
    bag = flat[leaves].tolist()
    weights = lens.astype(np.float64)
    target = np.random.choice(n, size=(len(bag),), p=weights / weights.sum()).tolist()
    fill = np.diff(stay_off).tolist()                                      # current length of every noised sentence
    for token, k in zip(bag, target):
        noised[k].insert(random.randint(0, fill[k]), token)
        fill[k] += 1
    return noised


def rand_perm(sentences, p=0.15):
    """A random subset of the batch's token positions is shuffled among itself (data_util.py:56-74)."""
    flat, off = flatten(sentences)
    picked = np.flatnonzero(np.random.uniform(size=flat.shape[0]) < p)
    tokens = flat[picked].tolist()
    random.shuffle(tokens)
    flat[picked] = tokens
    return split(flat, off)


# -- device-resident counterparts (padded CUDA batches) ----------------------------------------------------------

def transfer_noise_cuda(x, p: float, pad_id: int = 0, generator=None, out_len=None):
    """`transfer_noise` for a padded CUDA batch `x` [B, L] (pad_id = PAD): returns a padded [B, out_len] tensor
    (default out_len = 2 L; a sentence that would outgrow it loses its last tokens, which needs > L arrivals).
    Nothing leaves the device and nothing synchronises with the host.  A moved token picks its target sentence
    with probability proportional to the original lengths and a uniformly random slot between the target's
    remaining tokens; the draws come from `generator` (torch), not from the numpy / random streams."""
    import torch
    B, L = x.shape
    W = int(out_len) if out_len is not None else 2 * L
    dev = x.device
    real = x != pad_id
    lens = real.sum(1)
    u = torch.rand((B, L), device=dev, generator=generator)
    leaves = real & (u < p)
    stays = real & ~leaves
    # sort key = target sentence + a fractional slot: staying token j of a sentence sits at (j + 0.5) / (L + 1),
    # an arriving token anywhere in [0, 1 - eps) of the target's occupied range
    nstay = stays.sum(1)
    slot_stay = (torch.cumsum(stays, 1) - 0.5)
    w = lens.to(torch.float64)
    target = torch.multinomial(w / w.sum().clamp_min(1.0) + (w.sum() == 0), B * L, replacement=True, generator=generator).view(B, L)
    slot_move = torch.rand((B, L), device=dev, generator=generator, dtype=torch.float64) * nstay[target].to(torch.float64)
    rows = torch.arange(B, device=dev).view(B, 1).expand(B, L)
    sent = torch.where(leaves, target, rows)
    slot = torch.where(leaves, slot_move, slot_stay.to(torch.float64))
    key = torch.where(real, sent.to(torch.float64) * (2.0 * L + 2.0) + slot, torch.full_like(slot, float("inf")))
    order = torch.argsort(key.view(-1))
    s_sent = sent.view(-1)[order]
    s_tok = x.reshape(-1)[order]
    s_real = real.view(-1)[order]
    # position inside the target sentence = rank among the tokens of the same sentence (masks, no boolean indexing:
    # that would read a count back to the host)
    counts = torch.zeros(B, dtype=torch.long, device=dev).scatter_add_(0, s_sent, s_real.to(torch.long))
    start = torch.cumsum(counts, 0) - counts
    pos = torch.arange(B * L, device=dev) - start[s_sent]
    ok = s_real & (pos < W)
    flat = torch.full((B * W + 1,), pad_id, dtype=x.dtype, device=dev)           # the last slot swallows pads / overflow
    flat.scatter_(0, torch.where(ok, s_sent * W + pos, torch.full_like(pos, B * W)), s_tok)
    flat[B * W] = pad_id
    return flat[:B * W].view(B, W)


def rand_perm_cuda(x, p: float = 0.15, pad_id: int = 0, generator=None):
    """`rand_perm` for a padded CUDA batch: a random subset of the real token positions is permuted among itself
    (two sorts, no host round trip: picked positions in index order receive the picked tokens in random order, every
    other position maps to itself)."""
    import torch
    n = x.numel()
    real = (x != pad_id).view(-1)
    picked = real & (torch.rand(n, device=x.device, generator=generator) < p)
    idx = torch.arange(n, device=x.device, dtype=torch.float64)
    natural = torch.argsort(torch.where(picked, idx, idx + n))
    shuffled = torch.argsort(torch.where(picked, torch.rand(n, device=x.device, generator=generator, dtype=torch.float64), idx + 2.0))
    out = torch.empty_like(x).view(-1)
    out[natural] = x.reshape(-1)[shuffled]
    return out.view_as(x)
