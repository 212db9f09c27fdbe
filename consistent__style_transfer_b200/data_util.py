"""Host-side list munging of the pretrain label pipeline: the drop-in for the three functions of the
reference's src/data_util.py that feed the WMD label (`align`, `transfer_noise`, `rand_perm`;
/root/reference/src/data_util.py:25-74).

These run on the CPU in the reference too (they are python list edits driven by the global numpy and
`random` generators, not arithmetic); they are restated here so that `collate_pretrain`
(loader.py) can run without the reference checkout, with the same draws in the same order -- given
the same `np.random.seed` / `random.seed` the outputs are identical to the reference's
(tests/golden/noise_cases.json.gz was produced by executing the reference's own source).
One fix: the reference's `np.float` (data_util.py:44) no longer exists in numpy >= 1.24; it meant
the builtin float, i.e. float64.
"""
from __future__ import annotations

import random

import numpy as np


def align(sentences, pad_value, max_len=None):
    # data_util.py:25-30
    if max_len is None:
        max_len = max([len(sent) for sent in sentences])
    lengths = [len(sent[:max_len]) for sent in sentences]
    sentences = [sent[:max_len] + [pad_value] * (max_len - len(sent)) for sent in sentences]
    return sentences, lengths, max_len


def transfer_noise(sentences, p):
    # data_util.py:32-54: every token leaves its sentence with probability p and lands at a random
    # position of a sentence drawn with probability proportional to the ORIGINAL sentence lengths
    word_bag, sentences_noise, lens = [], [], []
    for s in sentences:
        s_noise = []
        ind = (np.random.uniform(size=(len(s))) < p)
        lens.append(len(s))
        for idx, v in enumerate(ind):
            if v:
                word_bag.append(s[idx])
            else:
                s_noise.append(s[idx])
        sentences_noise.append(s_noise)
    lens = np.array(lens, dtype=np.float64)
    p = lens / lens.sum()
    indexes = list(range(len(p)))
    choices = np.random.choice(indexes, size=(len(word_bag),), p=p)
    for idx, w in enumerate(word_bag):
        index = choices[idx]
        pos = random.randint(0, len(sentences_noise[index]))
        sentences_noise[index].insert(pos, w)
    return sentences_noise


def rand_perm(sentences, p=0.15):
    # data_util.py:56-74: a random subset of all token positions of the batch is shuffled among itself
    sent_lens, long_seq = [], []
    for sentence in sentences:
        long_seq += sentence
        sent_lens.append(len(sentence))
    ind = (np.random.uniform(size=(len(long_seq))) < p)
    hint_ids, words = [], []
    for idx, v in enumerate(ind):
        if v:
            hint_ids.append(idx)
            words.append(long_seq[idx])
    random.shuffle(words)
    for idx, id_ in enumerate(hint_ids):
        long_seq[id_] = words[idx]
    sentences, end_idx = [], 0
    for sent_len in sent_lens:
        sentences.append(long_seq[end_idx: end_idx + sent_len])
        end_idx += sent_len
    return sentences
