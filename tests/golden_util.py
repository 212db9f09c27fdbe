"""Loads the committed fixtures of tests/golden/ (made by tests/golden/make_golden.py)."""
import base64
import gzip
import json
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def text_cases():
    with gzip.open(os.path.join(GOLDEN, "wmd_text_cases.json.gz"), "rt", encoding="utf-8") as f:
        cases = json.load(f)
    for c in cases:
        raw = np.frombuffer(base64.b64decode(c["raw_vectors_int8_b64"]), dtype=np.int8)
        c["raw_vectors"] = raw.reshape(len(c["vocab"]), c["d"]).astype(np.float32)
        c["wmd"] = np.array([float.fromhex(h) for h in c["wmd_hex"]], dtype=np.float64)
    return cases


def tokenizer_cases():
    with gzip.open(os.path.join(GOLDEN, "tokenizer_cases.json.gz"), "rt", encoding="utf-8") as f:
        return json.load(f)


def pyemd_known_answers():
    with open(os.path.join(GOLDEN, "pyemd_known_answers.json")) as f:
        return json.load(f)


def same_floats(a, b):
    """bit-exact equality of two float64 arrays (inf == inf, 0.0 == 0.0)."""
    a = np.asarray(a, np.float64); b = np.asarray(b, np.float64)
    return a.shape == b.shape and a.tobytes() == b.tobytes()


def noise_cases():
    with gzip.open(os.path.join(GOLDEN, "noise_cases.json.gz"), "rt", encoding="utf-8") as f:
        return json.load(f)
