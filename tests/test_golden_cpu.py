"""CPU checks against the committed fixtures: the oracle still says what the fixtures froze, the
word tokenizer reproduces the reference tokenizer's tokens, and the artefact readers round-trip."""
import os
import pickle
import sys
import types

import numpy as np
import pytest

from golden_util import pyemd_known_answers, same_floats, text_cases, tokenizer_cases


@pytest.fixture(scope="module")
def cases():
    return text_cases()


def test_pyemd_known_answers(oracle):
    for k in pyemd_known_answers():
        got = oracle.emd(k["p"], k["q"], np.array(k["D"], float), extra_mass_penalty=k["penalty"])
        assert round(got, k["decimals"]) == round(k["want"], k["decimals"]), (k, got)


def test_oracle_python_loop_reproduces_golden(oracle, cases):
    for c in cases:
        kv = oracle.KeyedVectorsOracle(c["vocab"], c["raw_vectors"], normalize=True)
        step = 7 if len(c["wmd"]) > 500 else 3                  # a strided subset keeps the python loop short
        idx = range(0, len(c["wmd"]), step)
        got = [kv.wmdistance(c["tokens1"][i], c["tokens2"][i]) for i in idx]
        assert same_floats(got, c["wmd"][list(idx)]), c["name"]


def test_oracle_c_batch_reproduces_golden(oracle, cases):
    from consistent__style_transfer_b200.workload import to_csr
    for c in cases:
        table = oracle.init_sims_replace(c["raw_vectors"])
        rank = oracle.string_rank(c["vocab"])
        ids1, off1 = to_csr(c["rows1"]); ids2, off2 = to_csr(c["rows2"])
        got, _ = oracle.batch_wmd(table, ids1, off1, ids2, off2, rank=rank, nthreads=4)
        assert same_floats(got, c["wmd"]), c["name"]


def test_word_tokenizer_matches_reference_tokens():
    from consistent__style_transfer_b200.text_tokenizer import tokenize
    for k in tokenizer_cases():
        assert tokenize(k["text"]) == k["tokens"], k["text"]


def test_golden_rows_are_the_tokens(cases):
    for c in cases:
        tok_id = {w: i for i, w in enumerate(c["vocab"])}
        oov = set(c["oov_tokens"])
        for toks, rows in zip(c["tokens1"][:50], c["rows1"][:50]):
            assert [tok_id.get(t, -1) for t in toks] == rows
            assert all((t in oov) == (r < 0) for t, r in zip(toks, rows))


# ---- artefact readers (host logic, no GPU) ---------------------------------------------------------
def _fake_gensim_pickle(path, words, vectors, big=False):
    """Lays a file out the way gensim's SaveLoad.save does [recalled]: a pickle of objects whose
    classes live under gensim.*; with big=True the table goes to <path>.wv.vectors.npy."""
    mods = {}
    for name in ("gensim", "gensim.models", "gensim.models.word2vec", "gensim.models.keyedvectors"):
        mods[name] = types.ModuleType(name)
    W2V = type("Word2Vec", (), {"__module__": "gensim.models.word2vec"})
    KV = type("Word2VecKeyedVectors", (), {"__module__": "gensim.models.keyedvectors"})
    Vocab = type("Vocab", (), {"__module__": "gensim.models.keyedvectors"})
    mods["gensim.models.word2vec"].Word2Vec = W2V
    mods["gensim.models.keyedvectors"].Word2VecKeyedVectors = KV
    mods["gensim.models.keyedvectors"].Vocab = Vocab
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        wv = KV()
        wv.index2word = list(words)
        wv.vector_size = vectors.shape[1]
        wv.vocab = {}
        for i, w in enumerate(words):
            v = Vocab(); v.index = i; v.count = 5
            wv.vocab[w] = v
        if big:
            wv.vectors = None
            wv.__dict__["__numpys"] = ["vectors"]
            np.save(path + ".wv.vectors.npy", vectors)
        else:
            wv.vectors = vectors
        m = W2V(); m.wv = wv; m.window = 5; m.min_count = 5
        with open(path, "wb") as f:
            pickle.dump(m, f, protocol=2)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


@pytest.mark.parametrize("big", [False, True])
def test_gensim_pickle_reader_needs_no_gensim(tmp_path, big):
    from consistent__style_transfer_b200 import gensim_pickle
    rng = np.random.default_rng(0)
    words = ["w%d" % i for i in range(40)]
    vec = rng.standard_normal((40, 10)).astype(np.float32)
    p = str(tmp_path / "m.bin")
    _fake_gensim_pickle(p, words, vec, big=big)
    assert "gensim" not in sys.modules
    got_w, got_v = gensim_pickle.read(p)
    assert got_w == words and got_v.tobytes() == vec.tobytes()


def test_gensim_pickle_reader_refuses_code(tmp_path):
    from consistent__style_transfer_b200 import gensim_pickle
    p = str(tmp_path / "evil.bin")
    with open(p, "wb") as f:
        pickle.dump(os.getcwd, f)
    with pytest.raises(pickle.UnpicklingError):
        gensim_pickle.read(p)


@pytest.mark.parametrize("binary", [True, False])
def test_word2vec_format_round_trip(tmp_path, binary):
    from consistent__style_transfer_b200 import gensim_pickle
    rng = np.random.default_rng(1)
    words = ["the", "MASK", "é", "a-b", "!!!"]
    vec = rng.standard_normal((5, 7)).astype(np.float32)
    p = str(tmp_path / "v.w2v")
    gensim_pickle.write_word2vec_format(p, words, vec, binary=binary)
    got_w, got_v = gensim_pickle.read_word2vec_format(p)
    assert got_w == words and got_v.tobytes() == vec.tobytes()


def test_own_vectors_file_round_trip(tmp_path):
    from consistent__style_transfer_b200 import wmd
    rng = np.random.default_rng(2)
    words = ["a", "b", "c"]
    vec = rng.standard_normal((3, 4)).astype(np.float32)
    p = str(tmp_path / "own.bin")
    wmd.save_vectors(p, words, vec)
    got_w, got_v = wmd.load_vectors(p)
    assert got_w == words and got_v.tobytes() == vec.tobytes()


def test_string_rank_matches_python_sort():
    from consistent__style_transfer_b200.wmd import string_rank
    words = ["b", "a", "MASK", "é", "Z", "!", "aa"]
    r = string_rank(words)
    assert [words[i] for i in np.argsort(r)] == sorted(words)
