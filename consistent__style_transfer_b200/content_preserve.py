"""Drop-in for the WMD half of the reference's CP metric (evaluate/auto/content_preserve.py).

``load_word2vec_model`` (content_preserve.py:38-41) and ``calculate_wmd_scores`` (:43-50, called
from evaluate/eval.py:42) keep their names, arguments and return types; the per-pair python loop
over ``wmd_model.wv.wmdistance`` becomes one batched call into libwmd_b200.so.  Like the reference
there is no ``inf`` guard: a pair with an empty side (after OOV removal) scores ``inf`` and the
caller's mean propagates it.  ``mask_style_words`` (:13-28) is the caller-side text step.
"""
from __future__ import annotations

from typing import Callable, Iterable, List, Optional, Sequence

from .text_tokenizer import tokenize as _tokenize
from .wmd import KeyedVectors, _Model, load_vectors

CUSTOM_STYLE = "MASK"


def mask_style_words(texts: Iterable[str], lexicon, tokenize: Callable[[str], List[str]] = _tokenize) -> List[str]:
    edited = []
    for text in texts:
        edited.append(" ".join(CUSTOM_STYLE if tok.lower() in lexicon else tok for tok in tokenize(text)))
    return edited


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
