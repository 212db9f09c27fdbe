"""GPU parity of the drop-in Python surface (consistent__style_transfer_b200.wmd / .content_preserve)
against the committed golden fixtures and against the oracle's restatement of src/wmd.py."""
import math

import numpy as np
import pytest

from golden_util import same_floats, text_cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cases():
    return text_cases()


class FakeBPE:
    """Shape of src/vocab.py:BPETokenizer as the path uses it: ids_to_tokens / __len__ / tokenizer.id_to_token."""

    def __init__(self, tokens):
        self.tokens = list(tokens)
        self.tokenizer = self

    def id_to_token(self, i):
        return self.tokens[i] if 0 <= i < len(self.tokens) else None

    def ids_to_tokens(self, ids):
        return [self.id_to_token(i) for i in ids]

    def __len__(self):
        return len(self.tokens)


def test_calculate_wmd_scores_matches_golden(cases):
    from consistent__style_transfer_b200 import content_preserve as cp
    for c in cases:
        model = cp.model_from_embeddings(c["vocab"], c["raw_vectors"], normalize=True)
        got = cp.calculate_wmd_scores(c["text1"], c["text2"], model)
        assert same_floats(got, c["wmd"]), c["name"]
        # the single-pair gensim-shaped call (content_preserve.py:47)
        for i in (0, 1, len(got) // 2, len(got) - 1):
            v = model.wv.wmdistance(c["tokens1"][i], c["tokens2"][i])
            assert isinstance(v, float) and same_floats([v], [c["wmd"][i]])
        model.wv.close()


def test_distance_table_keeps_the_golden_values(cases):
    # KeyedVectors.enable_distance_table: looked-up cost tiles, same golden bits (real text of the reference)
    from consistent__style_transfer_b200 import content_preserve as cp
    for c in cases:
        model = cp.model_from_embeddings(c["vocab"], c["raw_vectors"], normalize=True)
        model.wv.enable_distance_table()
        got = cp.calculate_wmd_scores(c["text1"], c["text2"], model)
        assert same_floats(got, c["wmd"]), c["name"]
        model.wv.close()


def test_load_word2vec_model_from_files(tmp_path, cases):
    from consistent__style_transfer_b200 import content_preserve as cp, gensim_pickle, wmd
    c = cases[2]
    p1 = str(tmp_path / "own.bin"); wmd.save_vectors(p1, c["vocab"], c["raw_vectors"])
    p2 = str(tmp_path / "w2v.bin"); gensim_pickle.write_word2vec_format(p2, c["vocab"], c["raw_vectors"], binary=True)
    for p in (p1, p2):
        model = cp.load_word2vec_model(p)
        got = cp.calculate_wmd_scores(c["text1"], c["text2"], model)
        assert same_floats(got, c["wmd"])
        model.wv.close()


def test_wmddistance_surface_and_label_fallbacks(oracle, cases):
    from consistent__style_transfer_b200.wmd import WMDdistance
    c = cases[0]
    w = WMDdistance.from_embeddings(c["vocab"], c["raw_vectors"], normalize=True)
    # a tokenizer whose ids are a permutation of the rows plus specials the table does not know
    rng = np.random.default_rng(5)
    V = len(c["vocab"])
    perm = rng.permutation(V)
    toks = ["<pad>", "<s>", "</s>", "<unk>"] + [c["vocab"][r] for r in perm]
    bpe = FakeBPE(toks)
    row_to_id = np.empty(V, np.int64); row_to_id[perm] = np.arange(V) + 4
    enc = lambda rows: [int(row_to_id[r]) if r >= 0 else 3 for r in rows]          # OOV -> <unk>
    xs1 = [enc(r) for r in c["rows1"][:300]]
    xs2 = [enc(r) for r in c["rows2"][:300]]
    xs1 += [[], [5, 6], [3, 3], [], [len(toks) + 7]]
    xs2 += [[5], [], [7], [], [5]]
    got = w.cal_wmd_label(xs1, xs2, bpe)
    kv = oracle.KeyedVectorsOracle(c["vocab"], c["raw_vectors"], normalize=True)
    want = oracle.WMDdistanceOracle(kv).cal_wmd_label(xs1, xs2, bpe)
    assert isinstance(got, list) and all(isinstance(v, float) for v in got)
    assert same_floats(got, want)
    assert got[300:304] == [1.0, 2.0, 1.5, 0.0] and got[304] == 1.0
    assert same_floats(got[:300], c["wmd"][:300])
    # cal_wmd on token strings (src/wmd.py:31-32)
    assert same_floats([w.cal_wmd(c["tokens1"][3], c["tokens2"][3])], [c["wmd"][3]])
    assert w.cal_wmd(["<unk>"], c["tokens2"][3]) == math.inf
    # save / load round trip.  Like gensim, save() after load() writes the already-normalised rows and
    # load() runs init_sims(replace=True) on them again (src/wmd.py:47-55): a second float32
    # normalisation, which is not the identity -- the oracle has to do the same to agree bit for bit.
    import os, tempfile
    with tempfile.TemporaryDirectory() as td:
        p = os.path.join(td, "yelp-w2v.bin")
        w.save(p)
        w2 = WMDdistance.load(p)
        kv2 = oracle.KeyedVectorsOracle(c["vocab"], oracle.init_sims_replace(c["raw_vectors"]), normalize=True)
        want2 = oracle.WMDdistanceOracle(kv2).cal_wmd_label(xs1, xs2, bpe)
        assert same_floats(w2.cal_wmd_label(xs1, xs2, bpe), want2)
        w2.model.wv.close()
    # padded CUDA-tensor hook: no host sync, PAD_ID = 0 skipped
    import torch
    L = max(max(map(len, xs1)), max(map(len, xs2)))
    A = torch.zeros((len(xs1), L), dtype=torch.long); B = torch.zeros((len(xs1), L), dtype=torch.long)
    for i, (a, b) in enumerate(zip(xs1, xs2)):
        A[i, :len(a)] = torch.tensor(a, dtype=torch.long); B[i, :len(b)] = torch.tensor(b, dtype=torch.long)
    out = w.cal_wmd_padded(A.cuda(), B.cuda(), bpe, pad_id=0)
    torch.cuda.synchronize()
    raw = [kv.wmdistance(bpe.ids_to_tokens(a), bpe.ids_to_tokens(b)) for a, b in zip(xs1, xs2)]
    assert same_floats(out.cpu().numpy(), raw)
    w.model.wv.close()


def test_non_lazy_constructor_is_loud_without_gensim():
    from consistent__style_transfer_b200.wmd import WMDdistance
    try:
        import gensim  # noqa: F401
        pytest.skip("gensim present")
    except ImportError:
        pass
    with pytest.raises(RuntimeError, match="gensim"):
        WMDdistance(["nope.txt"], None)


def test_collate_pretrain_matches_reference_pipeline(oracle, cases):
    """src/loader.py:46-70 end to end: the reference's noising (frozen from its own source in
    tests/golden/noise_cases.json.gz) + the batched WMD label against the oracle's per-pair loop."""
    import random

    import torch
    from golden_util import noise_cases
    from consistent__style_transfer_b200 import data_util
    from consistent__style_transfer_b200.loader import collate_pretrain
    from consistent__style_transfer_b200.wmd import WMDdistance
    c = cases[0]
    w = WMDdistance.from_embeddings(c["vocab"], c["raw_vectors"], normalize=True)
    V = len(c["vocab"])
    toks = ["<pad>", "<s>", "</s>", "<unk>"] + list(c["vocab"])
    bpe = FakeBPE(toks)
    kv = oracle.KeyedVectorsOracle(c["vocab"], c["raw_vectors"], normalize=True)
    ow = oracle.WMDdistanceOracle(kv)
    for nc in noise_cases()[:3]:
        # golden batches hold ids >= 4 from a larger word list: fold them into this table's id range
        batch = [[4 + (t % V) for t in s] for s in nc["batch"]]
        samples = [(s, i % 2) for i, s in enumerate(batch)]
        np.random.seed(nc["seed"]); random.seed(nc["seed"] + 1000)
        out = collate_pretrain(bpe, w)(samples)
        assert len(out) == 6 and [t.dtype for t in out] == [torch.long] * 5 + [torch.float32]
        # replay the same draws to recover the noised sentences the labels were computed on
        np.random.seed(nc["seed"]); random.seed(nc["seed"] + 1000)
        n1 = data_util.transfer_noise([list(s) for s in batch], p=0.15)
        n2 = data_util.transfer_noise([list(s) for s in batch], p=0.15)
        want = ow.cal_wmd_label(n1, n2, bpe)
        assert out[5].numpy().tobytes() == np.asarray(want, np.float32).tobytes()
        assert out[0].shape[0] == len(batch) and out[1].shape[0] == len(batch)
        assert out[4].tolist() == [i % 2 for i in range(len(batch))]
        # the four id matrices are the padded sentences themselves (loader.py:54-58)
        assert out[0].tolist() == data_util.align(batch, 0)[0] and out[1].tolist() == data_util.align(n1, 0)[0]
        assert out[2].tolist() == data_util.align(n2, 0)[0]
    w.model.wv.close()


def test_async_labels_prefetcher_and_device_collate(oracle, cases):
    """cal_wmd_label_async == cal_wmd_label; LabelPrefetcher yields what collate_pretrain yields with the labels
    of the next batch in flight; collate_pretrain_cuda's labels are the WMD of ITS noised tensors."""
    import random

    import torch
    from golden_util import noise_cases
    from consistent__style_transfer_b200.loader import LabelPrefetcher, collate_pretrain, collate_pretrain_cuda
    from consistent__style_transfer_b200.wmd import WMDdistance
    c = cases[0]
    w = WMDdistance.from_embeddings(c["vocab"], c["raw_vectors"], normalize=True)
    V = len(c["vocab"])
    bpe = FakeBPE(["<pad>", "<s>", "</s>", "<unk>"] + list(c["vocab"]))
    batches = []
    for nc in noise_cases()[:3]:
        batch = [[4 + (t % V) for t in s] for s in nc["batch"]]
        batches.append([(s, i % 2) for i, s in enumerate(batch)])
    xs1 = [s for s, _ in batches[0]]; xs2 = [s[::-1][:-1] for s, _ in batches[0]]
    p = w.cal_wmd_label_async(xs1, xs2, bpe)
    with pytest.raises(RuntimeError, match="in flight"):
        w.model.wv.wmdistance(["a"], ["b"])                     # one job per handle
    assert same_floats(p.result(), w.cal_wmd_label(xs1, xs2, bpe))
    assert p.tensor(torch.float).dtype == torch.float32
    assert w.cal_wmd_label_async([], [], bpe).result() == []
    np.random.seed(3); random.seed(4)
    seq = [collate_pretrain(bpe, w)(b) for b in batches]
    np.random.seed(3); random.seed(4)
    pre = list(LabelPrefetcher(batches, bpe, w))
    assert len(pre) == len(seq)
    for a, b in zip(seq, pre):
        assert all(torch.equal(x, y) for x, y in zip(a, b))
    g = torch.Generator(device="cuda").manual_seed(5)
    out = collate_pretrain_cuda(bpe, w, generator=g)(batches[0])
    assert all(t.is_cuda for t in out) and out[5].dtype == torch.float32 and out[1].shape[0] == len(batches[0])
    n1 = [[t for t in row if t != 0] for row in out[1].tolist()]
    n2 = [[t for t in row if t != 0] for row in out[2].tolist()]
    want = w.cal_wmd_label(n1, n2, bpe)
    assert out[5].cpu().numpy().tobytes() == np.asarray(want, np.float32).tobytes()
    assert sorted(t for s in n1 for t in s) == sorted(t for s, _ in batches[0] for t in s)
    w.model.wv.close()
