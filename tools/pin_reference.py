#!/usr/bin/env python
"""Pin the oracle (and with it every parity claim of this repo) to the reference's REAL stack.

The arithmetic of the WMD path lives in gensim 3.8.x (`KeyedVectors.wmdistance`) and pyemd 0.5.1 (`emd`), which
are neither vendored by the reference nor installable in the build container; the fixtures under
tests/golden/ therefore freeze the ORACLE's output and DESIGN.md says "parity unpinned".  This script closes that
gap on any machine where the two packages import:

    pip install "gensim==3.8.3" "pyemd==0.5.1"          # python <= 3.8 wheels exist
    python tools/pin_reference.py --reference /path/to/consistent__style_transfer

It runs, through the reference's OWN modules wherever they import,
  1. `evaluate/auto/content_preserve.py:calculate_wmd_scores` and `src/wmd.py:WMDdistance.cal_wmd_label` on the
     golden inputs (tests/golden/wmd_text_cases.json.gz: the reference's shipped sentences, its own tokenizer,
     a seeded table normalised by `init_sims(replace=True)`) and diffs every value against the frozen IEEE hex
     with tolerance 0;
  2. `pyemd.emd` on tests/golden/pyemd_known_answers.json;
  3. `src/data_util.py:transfer_noise / rand_perm / align` on tests/golden/noise_cases.json.gz;
  4. a real `Word2Vec.save` pickle through `consistent__style_transfer_b200.gensim_pickle.read` (the reader's
     layout is recalled from gensim 3.8's `utils.SaveLoad`; until this step has run it is UNVERIFIED).
Prints one JSON report; exit code 0 only if every check that could run matched bit for bit.  Without gensim / pyemd
it reports `"status": "unavailable"` and exits 2 -- nothing is pinned by a run that could not import them.
"""
from __future__ import annotations

import argparse
import importlib
import json
import os
import random
import struct
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def bits(x: float) -> bytes:
    return struct.pack("<d", float(x))


def diff_report(got, want_hex):
    want = [float.fromhex(h) for h in want_hex]
    bad = [i for i, (g, w) in enumerate(zip(got, want)) if bits(g) != bits(w)]
    fin = [(g, w) for g, w in zip(got, want) if np.isfinite(g) and np.isfinite(w) and w != 0.0]
    rel = max((abs(g - w) / abs(w) for g, w in fin), default=0.0)
    return {"n": len(want), "bit_mismatches": len(bad), "first_mismatches": bad[:5], "max_rel_diff": rel}


class FakeBPE:
    """Shape of src/vocab.py:BPETokenizer as src/wmd.py uses it (ids_to_tokens)."""

    def __init__(self, tokens):
        self.tokens = list(tokens)

    def ids_to_tokens(self, ids):
        return [self.tokens[i] if 0 <= i < len(self.tokens) else None for i in ids]


def reference_function(ref, rel_path, module_dirs, module_name, func_name, fallback_ns):
    """The function object from the reference's own module; if the module's unrelated imports fail (sklearn's removed
    joblib shim, pytorch_lightning, ...), the function's source is executed from the file where it lies."""
    for d in module_dirs:
        p = os.path.join(ref, d)
        if p not in sys.path:
            sys.path.insert(0, p)
    try:
        return getattr(importlib.import_module(module_name), func_name), "imported"
    except Exception as exc:                                   # noqa: BLE001 - any import problem of the surroundings
        src = open(os.path.join(ref, rel_path), encoding="utf-8").read()
        import ast
        tree = ast.parse(src)
        ns = dict(fallback_ns)
        for node in tree.body:
            if isinstance(node, (ast.FunctionDef, ast.ClassDef)) and node.name == func_name:
                exec(compile(ast.Module([node], []), rel_path, "exec"), ns)
                return ns[func_name], f"source of {rel_path} executed in place ({type(exc).__name__}: {exc})"
        raise


def reference_tokenize(ref):
    src = open(os.path.join(ref, "evaluate/auto/tokenizer.py"), encoding="utf-8").read()
    if sys.version_info >= (3, 11):                            # the mid-pattern (?i) of tokenizer.py:37 (see make_golden.py)
        src = src.replace("r'(?i)' + t", "t").replace("re.UNICODE)", "re.UNICODE | re.IGNORECASE)")
    ns = {}
    exec(compile(src, "evaluate/auto/tokenizer.py", "exec"), ns)
    return ns["tokenize"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reference", default="/root/reference")
    a = ap.parse_args()
    report = {"reference": a.reference, "checks": {}}
    try:
        import gensim
        import pyemd
        from gensim.models import KeyedVectors
        report["gensim"], report["pyemd"] = gensim.__version__, getattr(pyemd, "__version__", "?")
    except Exception as exc:                                   # noqa: BLE001
        report.update(status="unavailable", why=f"{type(exc).__name__}: {exc}",
                      note="parity stays UNPINNED: install gensim 3.8.x and pyemd 0.5.1 and run this script again")
        print(json.dumps(report, indent=1))
        return 2
    from golden_util import noise_cases, pyemd_known_answers, text_cases

    tokenize = reference_tokenize(a.reference)
    calc, how_calc = reference_function(a.reference, "evaluate/auto/content_preserve.py", ["evaluate"], "auto.content_preserve",
                                        "calculate_wmd_scores", {"tokenize": tokenize})
    WMDdistance, how_wmd = reference_function(a.reference, "src/wmd.py", ["src"], "wmd", "WMDdistance", {"os": os})
    report["how"] = {"calculate_wmd_scores": how_calc, "WMDdistance": how_wmd}

    # 1. the golden text cases through gensim's own wmdistance -> pyemd
    for c in text_cases():
        kv = KeyedVectors(vector_size=c["d"])
        kv.add(list(c["vocab"]), np.asarray(c["raw_vectors"], np.float32))
        kv.init_sims(replace=True)                             # src/wmd.py:54, content_preserve.py:40

        class Model:                                           # what load_word2vec_model returns: only .wv is used
            wv = kv

        got = calc(c["text1"], c["text2"], Model)
        report["checks"][f"calculate_wmd_scores/{c['name']}"] = diff_report(got, c["wmd_hex"])
        w = WMDdistance(None, None, lazy=True)
        w.model = Model
        toks = list(c["vocab"])
        enc = lambda rows: [r if r >= 0 else len(toks) + 3 for r in rows]          # OOV -> an id the tokenizer does not know
        n = min(300, len(c["rows1"]))
        labels = w.cal_wmd_label([enc(r) for r in c["rows1"][:n]], [enc(r) for r in c["rows2"][:n]], FakeBPE(toks))
        want = []
        for h, r1, r2 in zip(c["wmd_hex"][:n], c["rows1"][:n], c["rows2"][:n]):   # the two fall-backs of src/wmd.py:37-44
            v = float.fromhex(h)
            if not r1 or not r2:
                v = float(max(len(r1), len(r2)))
            elif v == float("inf"):
                v = (len(r1) + len(r2)) / 2
            want.append(v.hex())
        report["checks"][f"cal_wmd_label/{c['name']}"] = diff_report(labels, want)

    # 2. pyemd's own known answers (recalled from its test-suite; here they meet the real package)
    ka = []
    for k in pyemd_known_answers():                            # keys: p, q, D, penalty, want, decimals
        v = pyemd.emd(np.asarray(k["p"], np.float64), np.asarray(k["q"], np.float64), np.asarray(k["D"], np.float64),
                      extra_mass_penalty=float(k.get("penalty", -1.0)))
        ka.append(abs(v - k["want"]) <= 0.5 * 10.0 ** (-k.get("decimals", 5)))
    report["checks"]["pyemd_known_answers"] = {"n": len(ka), "bit_mismatches": int(sum(not x for x in ka))}

    # 3. the noising functions of src/data_util.py on the frozen batches
    src = open(os.path.join(a.reference, "src/data_util.py"), encoding="utf-8").read().replace("np.float)", "np.float64)")
    ns = {}
    try:
        exec(compile(src, "src/data_util.py", "exec"), ns)
        bad = 0
        cases = noise_cases()
        for c in cases:
            np.random.seed(c["seed"]); random.seed(c["seed"] + 1000)
            n1 = ns["transfer_noise"]([list(s) for s in c["batch"]], p=0.15)
            n2 = ns["transfer_noise"]([list(s) for s in c["batch"]], p=0.15)
            n3 = ns["rand_perm"]([list(s) for s in c["batch"]], p=0.15)
            bad += [[int(t) for t in s] for s in n1] != c["noise1"]
            bad += [[int(t) for t in s] for s in n2] != c["noise2"]
            bad += [[int(t) for t in s] for s in n3] != c["perm"]
        report["checks"]["data_util_noise"] = {"n": 3 * len(cases), "bit_mismatches": int(bad)}
    except Exception as exc:                                   # noqa: BLE001 (torch missing in a gensim-3.8 environment, ...)
        report["checks"]["data_util_noise"] = {"skipped": f"{type(exc).__name__}: {exc}"}

    # 4. a real gensim pickle through the gensim-free reader
    from gensim.models.word2vec import Word2Vec
    from consistent__style_transfer_b200 import gensim_pickle
    rng = np.random.default_rng(0)
    words = ["tok%03d" % i for i in range(200)]
    sents = [[words[int(t)] for t in rng.integers(0, 200, size=12)] for _ in range(400)]
    try:
        m = Word2Vec(sents, iter=2, min_count=1)
    except TypeError:
        m = Word2Vec(sents, epochs=2, min_count=1)
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "w2v.bin")
        m.save(p)
        got_words, got_vec = gensim_pickle.read(p)
    ref_words = list(getattr(m.wv, "index2word", None) or m.wv.index_to_key)
    ok = list(got_words) == ref_words and np.asarray(got_vec, np.float32).tobytes() == np.asarray(m.wv.vectors, np.float32).tobytes()
    report["checks"]["gensim_pickle_reader"] = {"n": 1, "bit_mismatches": 0 if ok else 1}

    failed = {k: v for k, v in report["checks"].items() if v.get("bit_mismatches")}
    report["status"] = "pinned" if not failed else "MISMATCH"
    report["failed"] = sorted(failed)
    print(json.dumps(report, indent=1, default=str))
    return 0 if not failed else 1


if __name__ == "__main__":
    sys.exit(main())
