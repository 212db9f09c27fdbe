"""CPU-side pieces of the CP front end (no device needed): lexicon file, style masking."""
import json

from consistent__style_transfer_b200 import content_preserve as cp


def test_load_lexicon_reads_the_reference_layout(tmp_path):
    # layout written by evaluate/auto/style_lexicon.py:88-95: {"binary sentiment": [[feature, weight], ...]}
    p = tmp_path / "lexicon_yelp.json"
    p.write_text(json.dumps({"binary sentiment": [["great", 2.5], ["awful", -3.1], ["great", 0.1]], "other": [["x", 1.0]]}), encoding="utf-8")
    assert cp.load_lexicon(str(p)) == {"great", "awful"}


def test_mask_style_words_follows_the_reference():
    # evaluate/auto/content_preserve.py:13-28: tokenise, compare lower-cased tokens, join with single spaces
    lex = {"great", "awful", ":)"}
    out = cp.mask_style_words(["The food was GREAT , service awful :)", "nothing to mask here"], lex)
    assert out == ["The food was MASK , service MASK MASK", "nothing to mask here"]
    assert cp.mask_style_words([], lex) == []
    assert cp.mask_style_words([""], lex) == [""]


def test_content_preservation_counts_inf_pairs():
    class FakeWV:
        def wmdistance_batch(self, d1, d2):
            return [float("inf") if not a or not b else float(abs(len(a) - len(b))) for a, b in zip(d1, d2)]

    class FakeModel:
        wv = FakeWV()

    r = cp.content_preservation(["a b c", "x", ""], ["a b", "x y z", "q"], {"b"}, FakeModel())
    assert r["n"] == 3 and r["n_inf"] == 1
    assert r["cp"] == float("inf")                     # what the reference's mean returns
    assert r["cp_finite"] == (1.0 + 2.0) / 2
    assert r["scores"][:2] == [1.0, 2.0]
