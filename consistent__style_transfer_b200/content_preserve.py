"""Drop-in for the WMD half of the reference's CP metric (evaluate/auto/content_preserve.py).

``load_word2vec_model`` (content_preserve.py:38-41) and ``calculate_wmd_scores`` (:43-50, called
from evaluate/eval.py:42) keep their names, arguments and return types; the per-pair python loop
over ``wmd_model.wv.wmdistance`` becomes one batched call into libwmd_b200.so.  Like the reference
there is no ``inf`` guard: a pair with an empty side (after OOV removal) scores ``inf`` and the
caller's mean propagates it.  ``mask_style_words`` (:13-28) is the caller-side text step and
``load_lexicon`` (evaluate/auto/style_lexicon.py:100-103) reads the style words it masks.
``content_preservation`` is new: the CP block of evaluate/eval.py:36-43 as one call that also says
how many pairs scored ``inf`` instead of letting them swallow the mean silently.
"""
from __future__ import annotations

import json
import math
from typing import Callable, Dict, Iterable, List, Optional, Sequence, Set

from .text_tokenizer import tokenize as _tokenize
from .wmd import KeyedVectors, _Model, load_vectors

CUSTOM_STYLE = "MASK"


def mask_style_words(texts: Iterable[str], lexicon, tokenize: Callable[[str], List[str]] = _tokenize) -> List[str]:
    edited = []
    for text in texts:
        edited.append(" ".join(CUSTOM_STYLE if tok.lower() in lexicon else tok for tok in tokenize(text)))
    return edited


def load_lexicon(lexicon_path: str, key: str = "binary sentiment") -> Set[str]:
    """Style words of a lexicon file written by the reference's ``train`` step: a JSON object whose
    ``"binary sentiment"`` entry lists ``[feature, weight]`` pairs (style_lexicon.py:88-103)."""
    with open(lexicon_path, "r", encoding="utf-8") as f:
        table = json.load(f)
    return {entry[0] for entry in table[key]}


def load_word2vec_model(path: str, device: int = 0) -> _Model:
    """Object with ``.wv.wmdistance``; vectors L2-normalised as ``init_sims(replace=True)`` does."""
    words, vectors = load_vectors(path)
    return _Model(KeyedVectors(words, vectors, normalize=True, device=device))


def model_from_embeddings(index2word: Sequence[str], vectors, normalize: bool = True, device: int = 0) -> _Model:
    return _Model(KeyedVectors(index2word, vectors, normalize=normalize, device=device))


def calculate_wmd_scores(references: Sequence[str], candidates: Sequence[str], wmd_model,
                         tokenize: Optional[Callable[[str], List[str]]] = None) -> List[float]:
    tok = tokenize or _tokenize
    n = len(references)
    docs1 = [tok(references[i]) for i in range(n)]
    docs2 = [tok(candidates[i]) for i in range(n)]
    return wmd_model.wv.wmdistance_batch(docs1, docs2)


def content_preservation(origin: Sequence[str], transfer: Sequence[str], lexicon, wmd_model,
                         tokenize: Optional[Callable[[str], List[str]]] = None) -> Dict[str, float]:
    """The CP metric as evaluate/eval.py:36-43 computes it -- style words of both sides masked, WMD of
    (masked transfer, masked origin), arithmetic mean -- with the ``inf`` pairs counted: ``cp`` is the
    reference's number (``inf`` as soon as one pair has a side without in-vocabulary tokens, exactly as
    ``sum(seq) / len(seq)`` gives), ``cp_finite`` the mean over the finite pairs, ``n_inf`` how many were not."""
    tok = tokenize or _tokenize
    masked_origin = mask_style_words(origin, lexicon, tok)
    masked_transfer = mask_style_words(transfer, lexicon, tok)
    scores = calculate_wmd_scores(masked_transfer, masked_origin, wmd_model, tok)
    finite = [x for x in scores if math.isfinite(x)]
    return {"cp": sum(scores) / len(scores) if scores else float("nan"),
            "cp_finite": sum(finite) / len(finite) if finite else float("nan"),
            "n_inf": len(scores) - len(finite), "n": len(scores), "scores": scores}
