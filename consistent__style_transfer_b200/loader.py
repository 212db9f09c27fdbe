"""`collate_pretrain`: the in-loop caller of the WMD path (reference src/loader.py:46-70, used by
src/main_pretrain.py:118-122), with the per-pair python loop of `cal_wmd_label` replaced by one
batched call into libwmd_b200.so.

Returns the same six tensors in the same order and dtypes (the WMD label ends up float32, as in
loader.py:68).  `w2v` is a `consistent__style_transfer_b200.wmd.WMDdistance`; it owns a CUDA handle,
so the collate function must run in the process that owns the device: the DataLoader default
`num_workers=0` of the reference (main_pretrain.py:120-122), or `multiprocessing_context="spawn"`.
"""
from __future__ import annotations

import torch

from .data_util import align, rand_perm, transfer_noise

PAD_ID = 0          # src/vocab.py:9


def pth_tensor(tensor, dtype):
    # data_util.py:15-16
    return torch.tensor(tensor, dtype=dtype)


def collate_pretrain(vocab, w2v):
    def collate_func(batch_samples):
        sentences, labels = zip(*batch_samples)

        noised_sentences_1 = transfer_noise(sentences, p=0.15)
        noised_sentences_2 = transfer_noise(sentences, p=0.15)
        noised_sentences_3 = rand_perm(sentences, p=0.15)

        aligned_sentences, _, _ = align(sentences, PAD_ID)
        aligned_noised_sentences_1, _, _ = align(noised_sentences_1, PAD_ID)
        aligned_noised_sentences_2, _, _ = align(noised_sentences_2, PAD_ID)

        aligned_noised_sentences_3, _, _ = align(noised_sentences_3, PAD_ID)

        c_label = w2v.cal_wmd_label(noised_sentences_1, noised_sentences_2, vocab)      # one GPU call per batch

        return (
            pth_tensor(aligned_sentences, torch.long),
            pth_tensor(aligned_noised_sentences_1, torch.long),
            pth_tensor(aligned_noised_sentences_2, torch.long),
            pth_tensor(aligned_noised_sentences_3, torch.long),
            pth_tensor(labels, torch.long),
            pth_tensor(c_label, torch.float)
        )
    return collate_func
