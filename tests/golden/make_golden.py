#!/usr/bin/env python
"""Generates the committed fixtures under tests/golden/ (run in the BUILD container, where
/root/reference exists; the tests never read /root/reference).

    python tests/golden/make_golden.py

What comes from where:
* text: the reference's own shipped sentences -- data/yelp/style.test.{0,1} paired line by line
  with the human rewrites in data/yelp/reference.{0,1} (the stand-in for BASELINE config 1, "Yelp
  test-set WMD between source and transferred sentences"), a slice of data/book/style.dev.*, and the
  style-masked (origin, transfer) cells of evaluate/user/result/yelp_0.csv (exactly the strings
  ``calculate_wmd_scores`` sees, evaluate/eval.py:40-42).
* tokens: produced by the REFERENCE's tokenizer source, evaluate/auto/tokenizer.py, executed from
  where it lies with its one Python-3.11+ incompatibility patched (the mid-pattern ``(?i)`` of
  tokenizer.py:37 is hoisted to ``re.IGNORECASE``, which is what Python <= 3.10 did with it).
* embedding table: the reference's trained word2vec models are not shipped
  (.MISSING_LARGE_BLOBS) -> a seeded random table, d = 100 (gensim's default size),
  one row per token; the raw rows (int8-valued, to keep the file small) are stored so that the L2
  normalisation done by ``load`` (init_sims(replace=True)) is part of the test.
* expected values: the reference CANNOT run here (gensim / pyemd absent), so the
  numbers are the CPU oracle's (oracle/wmd_oracle.py, the gensim-shaped per-pair python loop +
  the C emd_hat restatement), stored as IEEE-754 hex so that comparisons are bit-exact.
  PARITY UNPINNED: these fixtures freeze the oracle, not a run of gensim + pyemd.
* pyemd known-answer vectors: recalled from pyemd 0.5.1's test-suite.
"""
import base64
import csv
import gzip
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = "/root/reference"

from oracle import wmd_oracle  # noqa: E402


def reference_tokenizer():
    src = open(os.path.join(REF, "evaluate/auto/tokenizer.py"), encoding="utf-8").read()
    src = src.replace("r'(?i)' + t", "t").replace("re.UNICODE)", "re.UNICODE | re.IGNORECASE)")
    ns = {}
    exec(compile(src, "reference:evaluate/auto/tokenizer.py", "exec"), ns)
    return ns["tokenize"]


def read_lines(rel):
    with open(os.path.join(REF, rel), encoding="utf-8") as f:
        return [line.rstrip("\n") for line in f]


def hexf(x):
    return float(x).hex()


def build_case(name, pairs_text, tokenize, d, seed, lower=False):
    docs1 = [tokenize(a) for a, _ in pairs_text]
    docs2 = [tokenize(b) for _, b in pairs_text]
    words = sorted({t for d_ in docs1 + docs2 for t in d_})
    rng = np.random.default_rng(seed)
    rng.shuffle(words)                                   # row order != string order: exercises the rank table
    # ~3 % of the tokens are left out of the vocabulary (OOV path)
    oov = set(w for w in words if rng.random() < 0.03)
    vocab = [w for w in words if w not in oov]
    raw_i8 = rng.integers(-127, 128, size=(len(vocab), d), dtype=np.int8)      # small integers keep the fixture small
    raw = raw_i8.astype(np.float32)
    kv = wmd_oracle.KeyedVectorsOracle(vocab, raw, normalize=True)
    values = [kv.wmdistance(a, b) for a, b in zip(docs1, docs2)]
    tok_id = {w: i for i, w in enumerate(vocab)}
    enc = lambda doc: [tok_id.get(t, -1) for t in doc]
    return {
        "name": name, "d": d, "vocab": vocab, "oov_tokens": sorted(oov),
        "raw_vectors_int8_b64": base64.b64encode(raw_i8.tobytes()).decode("ascii"),
        "text1": [a for a, _ in pairs_text], "text2": [b for _, b in pairs_text],
        "rows1": [enc(x) for x in docs1], "rows2": [enc(x) for x in docs2],
        "tokens1": docs1, "tokens2": docs2,
        "wmd_hex": [hexf(v) for v in values],
    }


def main():
    tokenize = reference_tokenizer()
    cases = []

    # (1) Yelp test sentences vs their human rewrites: 1000 pairs
    pairs = []
    for lab in (0, 1):
        src = read_lines(f"data/yelp/style.test.{lab}")
        ref = [l.split("\t") for l in read_lines(f"data/yelp/reference.{lab}")]
        assert len(src) == len(ref)
        for s, r in zip(src, ref):
            assert r[0].strip() == s.strip()
            pairs.append((s, r[1]))
    cases.append(build_case("yelp_test_vs_human_rewrite", pairs, tokenize, d=100, seed=11))

    # (2) book dev sentences: neighbouring lines as (long) pairs, 400 pairs
    book = read_lines("data/book/style.dev.0")[:400] + read_lines("data/book/style.dev.1")[:400]
    pairs = [(book[i], book[i + 1]) for i in range(0, 800, 2)]
    cases.append(build_case("book_dev_neighbours", pairs, tokenize, d=100, seed=12))

    # (3) style-masked (origin, transfer) cells of the user study: what calculate_wmd_scores sees
    pairs = []
    with open(os.path.join(REF, "evaluate/user/result/yelp_0.csv"), encoding="utf-8") as f:
        origin = None
        for row in csv.DictReader(f):
            if row["origin"]:
                origin = row["origin"].split("\n")[-1]
            if row["transfer"] and origin:
                pairs.append((row["transfer"].split("\n")[-1], origin))     # eval.py:42 order: (transfer, origin)
    cases.append(build_case("yelp_user_study_masked", pairs[:300], tokenize, d=100, seed=13))

    with gzip.open(os.path.join(HERE, "wmd_text_cases.json.gz"), "wt", encoding="utf-8") as f:
        json.dump(cases, f)

    # tokenizer fixture: reference tokens for a few hundred shipped lines + hand-made hard cases
    hard = ["I can't believe it's MASK-free :) www.foo.com/x?y=1 #yay ##hi U.S.A. Mr. Smith's 12.5% ... !!!",
            "e-mail me@site.co.uk <3 <33 :-D xD ;P o_O ^_^ @user @@user $$$ a_b c-d  tabs\tand\nnewlines",
            "MR. JONES AND DR. WHO met prof. x :D :d =) ☀ ££ 100%%"]
    lines = read_lines("data/yelp/style.dev.0")[:150] + read_lines("data/book/style.dev.1")[:150] + \
        [r[1] for r in (l.split("\t") for l in read_lines("data/yelp/reference.1")[:100])] + hard
    with gzip.open(os.path.join(HERE, "tokenizer_cases.json.gz"), "wt", encoding="utf-8") as f:
        json.dump([{"text": t, "tokens": tokenize(t)} for t in lines], f)

    # pyemd known-answer vectors (recalled from pyemd 0.5.1 test/test_pyemd.py)
    kat = [
        {"p": [0.0, 1.0], "q": [5.0, 3.0], "D": [[0.0, 0.5], [0.5, 0.0]], "penalty": -1.0, "want": 3.5, "decimals": 5},
        {"p": [1.0, 1.0], "q": [1.0, 1.0], "D": [[0.0, 1.0], [1.0, 0.0]], "penalty": -1.0, "want": 0.0, "decimals": 5},
        {"p": [0.0, 2.0, 1.0, 2.0], "q": [2.0, 1.0, 2.0, 1.0],
         "D": [[0.0, 1.0, 1.0, 2.0], [1.0, 0.0, 2.0, 1.0], [1.0, 2.0, 0.0, 1.0], [2.0, 1.0, 1.0, 0.0]],
         "penalty": 2.5, "want": 4.5, "decimals": 5},
        {"p": [0.9, 0.1], "q": [0.2, 0.8], "D": [[1.0, 1.0], [1.0, 1.0]], "penalty": -1.0, "want": 0.7, "decimals": 5},
    ]
    with open(os.path.join(HERE, "pyemd_known_answers.json"), "w") as f:
        json.dump(kat, f, indent=1)
    for c in cases:
        v = np.array([float.fromhex(h) for h in c["wmd_hex"]])
        print(c["name"], len(v), "pairs; finite", int(np.isfinite(v).sum()), "mean", float(v[np.isfinite(v)].mean()),
              "zeros", int((v == 0).sum()), "V", len(c["vocab"]))


if __name__ == "__main__" and "--noise" not in sys.argv:
    main()


def make_noise_cases():
    """Runs the REFERENCE's own transfer_noise / rand_perm / align (src/data_util.py, executed from
    where it lies; its one numpy >= 1.24 incompatibility, `np.float`, is aliased to the builtin it
    meant) on seeded batches of shipped sentences encoded as word ids, and freezes inputs + outputs."""
    import random
    src = open(os.path.join(REF, "src/data_util.py"), encoding="utf-8").read()
    if not hasattr(np, "float"):
        np.float = float
    ns = {}
    exec(compile(src, "reference:src/data_util.py", "exec"), ns)
    lines = read_lines("data/yelp/style.dev.0")[:600] + read_lines("data/book/style.dev.0")[:300]
    vocab = {}
    sents = [[vocab.setdefault(w, len(vocab) + 4) for w in line.split()][:30] for line in lines if line.strip()]
    cases = []
    for k, (lo, hi, seed) in enumerate([(0, 256, 11), (256, 512, 12), (600, 728, 13), (40, 41, 14), (100, 103, 15)]):
        batch = [list(s) for s in sents[lo:hi]]
        np.random.seed(seed); random.seed(seed + 1000)
        n1 = ns["transfer_noise"]([list(s) for s in batch], p=0.15)
        n2 = ns["transfer_noise"]([list(s) for s in batch], p=0.15)
        n3 = ns["rand_perm"]([list(s) for s in batch], p=0.15)
        al, lens, ml = ns["align"]([list(s) for s in n1], 0)
        cases.append({"seed": seed, "batch": batch, "noise1": [[int(t) for t in s] for s in n1],
                      "noise2": [[int(t) for t in s] for s in n2], "perm": [[int(t) for t in s] for s in n3],
                      "aligned1": al, "lengths1": lens, "max_len1": ml})
    with gzip.open(os.path.join(HERE, "noise_cases.json.gz"), "wt", encoding="utf-8") as f:
        json.dump(cases, f)
    print("noise cases:", len(cases))


if __name__ == "__main__" and "--noise" in sys.argv:
    make_noise_cases()
