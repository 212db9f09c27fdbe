"""Pretrain batch assembly around the WMD label (the in-loop caller of the path).

`collate_pretrain(vocab, w2v)` stands where the reference's src/loader.py:46-70 stands (used by
src/main_pretrain.py:118-122) and returns the same six tensors in the same order and dtypes -- sentences, two
`transfer_noise` copies, one `rand_perm` copy (all int64, padded with PAD_ID), the style labels (int64) and the WMD
label between the two noised copies (float32, loader.py:68) -- from the same generator draws in the same order.
What is different is the schedule: the label's GPU work is queued (`WMDdistance.cal_wmd_label_async`) as soon as
the two noised copies exist, the four id matrices are padded straight into arrays while the kernels run, and the
label is collected last.  `w2v` owns a CUDA handle, so the collate function must run in the process that owns the
device: the DataLoader default `num_workers=0` of the reference, or `multiprocessing_context="spawn"`.

`collate_pretrain_cuda` keeps the whole batch on the device: one host-to-device copy of the padded sentences,
noising with `data_util.transfer_noise_cuda` / `rand_perm_cuda`, labels through the padded device entry of the
library (`WMDdistance.cal_wmd_padded`), nothing comes back to the host -- the step is then bound by the GPU, not
by python list editing.  Its draws come from a torch generator: same distribution as the reference, different stream.

`LabelPrefetcher` wraps any iterable of sample batches and keeps the NEXT batch's labels in flight while the
caller trains on the current one.
"""
from __future__ import annotations

import numpy as np
import torch

from . import data_util
from .data_util import align_array, rand_perm, transfer_noise

PAD_ID = 0          # src/vocab.py:9
NOISE_P = 0.15      # src/loader.py:50-52


def _long(grid: np.ndarray) -> torch.Tensor:
    return torch.from_numpy(grid)


def _start(batch_samples, vocab, w2v):
    """Noising + label submission of one batch; returns everything `_finish` needs."""
    sentences, labels = zip(*batch_samples)
    noised_1 = transfer_noise(sentences, p=NOISE_P)
    noised_2 = transfer_noise(sentences, p=NOISE_P)
    permuted = rand_perm(sentences, p=NOISE_P)
    pending = w2v.cal_wmd_label_async(noised_1, noised_2, vocab)          # the GPU works from here on
    return sentences, noised_1, noised_2, permuted, labels, pending


def _finish(state):
    sentences, noised_1, noised_2, permuted, labels, pending = state
    grids = [_long(align_array(s, PAD_ID)) for s in (sentences, noised_1, noised_2, permuted)]
    return (*grids, torch.tensor(labels, dtype=torch.long), pending.tensor(torch.float))


def collate_pretrain(vocab, w2v):
    def collate_func(batch_samples):
        return _finish(_start(batch_samples, vocab, w2v))
    return collate_func


class LabelPrefetcher:
    """Iterates over collated pretrain batches with the labels of batch k+1 computed while the consumer works on
    batch k: `for batch in LabelPrefetcher(sample_batches, vocab, w2v): train(batch)`."""

    def __init__(self, sample_batches, vocab, w2v):
        self._it, self._vocab, self._w2v = iter(sample_batches), vocab, w2v

    def __iter__(self):
        ahead = None
        for samples in self._it:
            if ahead is not None:
                done = _finish(ahead)                 # waits for the labels queued one iteration ago
                ahead = _start(samples, self._vocab, self._w2v)
                yield done
            else:
                ahead = _start(samples, self._vocab, self._w2v)
        if ahead is not None:
            yield _finish(ahead)


def collate_pretrain_cuda(vocab, w2v, device=None, generator=None):
    """Device-resident variant: returns the same six tensors, all on the GPU (label float32)."""
    dev = torch.device("cuda", w2v.device) if device is None else torch.device(device)

    def collate_func(batch_samples):
        sentences, labels = zip(*batch_samples)
        x = _long(align_array(sentences, PAD_ID)).to(dev, non_blocking=True)
        noised_1 = data_util.transfer_noise_cuda(x, NOISE_P, PAD_ID, generator)
        noised_2 = data_util.transfer_noise_cuda(x, NOISE_P, PAD_ID, generator)
        permuted = data_util.rand_perm_cuda(x, NOISE_P, PAD_ID, generator)
        dist = w2v.cal_wmd_padded(noised_1, noised_2, vocab, pad_id=PAD_ID)
        # the fall-backs of src/wmd.py:37-44 on the device: empty side -> max(len); inf -> mean of the lengths
        len1 = (noised_1 != PAD_ID).sum(1).to(torch.float64)
        len2 = (noised_2 != PAD_ID).sum(1).to(torch.float64)
        label = torch.where(torch.isinf(dist), (len1 + len2) / 2, dist)
        label = torch.where((len1 == 0) | (len2 == 0), torch.maximum(len1, len2), label)
        return (x, noised_1, noised_2, permuted, torch.tensor(labels, dtype=torch.long, device=dev), label.to(torch.float))
    return collate_func
