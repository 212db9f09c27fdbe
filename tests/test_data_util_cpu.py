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
