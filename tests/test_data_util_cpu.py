"""The pretrain label pipeline's list munging (consistent__style_transfer_b200/data_util.py) against
outputs of the REFERENCE's own src/data_util.py, frozen by tests/golden/make_golden.py --noise:
same seeds -> same draws in the same order -> identical noised sentences."""
import random

import numpy as np

from golden_util import noise_cases

from consistent__style_transfer_b200 import data_util


def test_transfer_noise_rand_perm_align_match_the_reference():
    cases = noise_cases()
    assert len(cases) >= 5
    for c in cases:
        batch = c["batch"]
        np.random.seed(c["seed"]); random.seed(c["seed"] + 1000)
        n1 = data_util.transfer_noise([list(s) for s in batch], p=0.15)
        n2 = data_util.transfer_noise([list(s) for s in batch], p=0.15)
        n3 = data_util.rand_perm([list(s) for s in batch], p=0.15)
        assert [[int(t) for t in s] for s in n1] == c["noise1"]
        assert [[int(t) for t in s] for s in n2] == c["noise2"]
        assert [[int(t) for t in s] for s in n3] == c["perm"]
        al, lens, ml = data_util.align([list(s) for s in n1], 0)
        assert al == c["aligned1"] and lens == c["lengths1"] and ml == c["max_len1"]
        # tokens are only moved between sentences, never created or lost
        assert sorted(t for s in n1 for t in s) == sorted(t for s in batch for t in s)
        assert [len(s) for s in n3] == [len(s) for s in batch]


def test_align_truncates_and_pads():
    s, lens, ml = data_util.align([[1, 2, 3], [4], []], 0, max_len=2)
    assert s == [[1, 2], [4, 0], [0, 0]] and lens == [2, 1, 0] and ml == 2


def test_flat_helpers_round_trip():
    docs = [[3, 4], [], [9], [1, 2, 3, 4, 5]]
    flat, off = data_util.flatten(docs)
    assert flat.tolist() == [3, 4, 9, 1, 2, 3, 4, 5] and off.tolist() == [0, 2, 2, 3, 8]
    assert data_util.split(flat, off) == docs
    g = data_util.align_array(docs, 0)
    assert g.shape == (4, 5) and g[0].tolist() == [3, 4, 0, 0, 0] and g[3].tolist() == [1, 2, 3, 4, 5]
    assert data_util.align_array(docs, 7, max_len=2).tolist() == [[3, 4], [7, 7], [9, 7], [1, 2]]
    pad, lens, ml = data_util.align(docs, 0)
    assert np.array_equal(np.asarray(pad), g) and lens == [2, 0, 1, 5] and ml == 5


def test_device_noising_on_cpu_tensors_moves_tokens_without_losing_any():
    """transfer_noise_cuda / rand_perm_cuda are plain tensor programs: on CPU tensors they must show the same
    invariants as the reference's list version (tokens are moved, never created or lost; rand_perm keeps every
    sentence's length), and the share of moved tokens must be about p."""
    import torch
    g = torch.Generator().manual_seed(7)
    B, L = 256, 18
    lens = torch.randint(1, L + 1, (B,), generator=g)
    x = torch.randint(4, 9000, (B, L), generator=g)
    x = torch.where(torch.arange(L).view(1, L) < lens.view(B, 1), x, torch.zeros_like(x))
    y = data_util.transfer_noise_cuda(x, 0.15, 0, g)
    assert y.shape == (B, 2 * L)
    assert sorted(x[x != 0].tolist()) == sorted(y[y != 0].tolist())
    # left-packed rows: no pad before a token
    real = y != 0
    assert bool(((real.cumsum(1) == torch.arange(1, 2 * L + 1).view(1, -1)) | ~real).all())
    moved = 0
    for b in range(B):
        a, c = x[b][x[b] != 0].tolist(), y[b][y[b] != 0].tolist()
        moved += len(a) - len(set(a) & set(c))
    frac = moved / int((x != 0).sum())
    assert 0.08 < frac < 0.22, frac
    z = data_util.rand_perm_cuda(x, 0.15, 0, g)
    assert torch.equal(z != 0, x != 0) and sorted(z[z != 0].tolist()) == sorted(x[x != 0].tolist())
    assert 0.03 < float((z != x).float().sum() / (x != 0).sum()) < 0.2
    # an all-pad batch is left alone
    e = torch.zeros((3, 5), dtype=torch.long)
    assert torch.equal(data_util.transfer_noise_cuda(e, 0.5, 0, g), torch.zeros((3, 10), dtype=torch.long))
