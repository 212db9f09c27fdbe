"""Read the embedding artefacts of the reference without gensim.

``WMDdistance.load`` (/root/reference/src/wmd.py:50-55) and ``load_word2vec_model``
(/root/reference/evaluate/auto/content_preserve.py:38-41) call ``Word2Vec.load(path)`` on a file
written by gensim's ``SaveLoad.save``: a pickle of the model object in which every class lives
under ``gensim.*``; numpy arrays above gensim's ``sep_limit`` (10 MiB) are not inside the pickle
but in side files ``<path>.<attr>.npy`` (``<path>.wv.vectors.npy`` for the table), their names
listed under ``__numpys`` in the owner's ``__dict__`` [recalled from gensim 3.8 utils.py; gensim is
not installed here, so this reader is exercised on pickles laid out that way by the tests].

Only two things are needed from the file: ``wv.index2word`` (gensim 4: ``index_to_key``) and
``wv.vectors`` (older: ``syn0``).  Every ``gensim.*`` class is therefore unpickled into an inert
attribute bag; nothing from the file is executed.
"""
from __future__ import annotations

import os
import pickle
import struct
from typing import List, Tuple

import numpy as np


class _Bag:
    """Inert stand-in for any gensim class."""

    def __init__(self, *a, **k):
        pass

    def __setstate__(self, state):
        if isinstance(state, dict):
            self.__dict__.update(state)
        elif isinstance(state, tuple) and len(state) == 2:              # (dict, slots)
            for part in state:
                if isinstance(part, dict):
                    self.__dict__.update(part)


_NUMPY_OK = {"_reconstruct", "ndarray", "dtype", "_frombuffer", "scalar"}
_BUILTIN_OK = {"set", "frozenset", "dict", "list", "tuple", "object", "int", "float", "str", "bytes", "bool",
               "complex", "slice", "range", "bytearray"}


class _StubUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        if module.split(".")[0] == "gensim":
            return type(name, (_Bag,), {"__module__": module})
        if module.split(".")[0] == "numpy" and name in _NUMPY_OK:
            return super().find_class(module, name)
        if module in ("numpy.random._pickle", "numpy.random.mtrand", "numpy.random"):
            return type(name, (_Bag,), {"__module__": module})         # RandomState inside the model: ignored
        if module in ("builtins", "__builtin__") and name in _BUILTIN_OK:
            return super().find_class("builtins", name)
        if module == "collections" and name in ("defaultdict", "OrderedDict"):
            return super().find_class(module, name)
        if module == "_codecs" and name == "encode":                       # protocol-2 numpy buffers (latin-1 text)
            return super().find_class(module, name)
        if module in ("copy_reg", "copyreg") and name == "_reconstructor":
            return _reconstructor
        raise pickle.UnpicklingError(f"refusing to unpickle {module}.{name}")


def _reconstructor(cls, base, state):
    return cls.__new__(cls)


def _side_array(path: str, owner_prefix: str, attr: str):
    f = f"{path}.{owner_prefix}{attr}.npy"
    return np.load(f, mmap_mode=None) if os.path.exists(f) else None


def read(path: str) -> Tuple[List[str], np.ndarray]:
    """(index2word, float32 [V, d]) from a gensim ``Word2Vec.save`` / ``KeyedVectors.save`` file."""
    with open(path, "rb") as f:
        model = _StubUnpickler(f, encoding="latin1").load()
    wv = getattr(model, "wv", model)                                    # a bare KeyedVectors has no .wv
    prefix = "wv." if wv is not model else ""
    vectors = None
    for attr in ("vectors", "syn0"):
        v = wv.__dict__.get(attr)
        if v is None:
            v = _side_array(path, prefix, attr)
        if v is not None:
            vectors = v
            break
    words = wv.__dict__.get("index2word") or wv.__dict__.get("index_to_key")
    if vectors is None or words is None:
        raise ValueError(f"{path}: no wv.vectors / wv.index2word found in the gensim pickle")
    vectors = np.ascontiguousarray(vectors, dtype=np.float32)
    words = [w if isinstance(w, str) else w.decode("utf-8") for w in words]
    if vectors.ndim != 2 or vectors.shape[0] != len(words):
        raise ValueError(f"{path}: {len(words)} tokens but a table of shape {vectors.shape}")
    return words, vectors


def read_word2vec_format(path: str) -> Tuple[List[str], np.ndarray]:
    """word2vec C format, text or binary (header line ``V d``)."""
    with open(path, "rb") as f:
        header = f.readline().split()
        if len(header) != 2:
            raise ValueError(f"{path}: not a word2vec-format file")
        V, d = int(header[0]), int(header[1])
        rest = f.read()
    words: List[str] = []
    vecs = np.empty((V, d), np.float32)
    first_line = rest.split(b"\n", 1)[0].split(b" ")
    as_text = False
    if len([x for x in first_line if x]) == d + 1:
        try:
            [float(x) for x in first_line[1:] if x]
            as_text = True
        except ValueError:
            as_text = False
    if as_text:
        lines = rest.decode("utf-8").splitlines()
        for i in range(V):
            parts = lines[i].rstrip().split(" ")
            words.append(parts[0])
            vecs[i] = np.asarray(parts[1:d + 1], dtype=np.float32)
        return words, vecs
    pos = 0
    for i in range(V):
        sp = rest.index(b" ", pos)
        words.append(rest[pos:sp].lstrip(b"\n").decode("utf-8"))
        pos = sp + 1
        vecs[i] = np.frombuffer(rest, dtype="<f4", count=d, offset=pos)
        pos += 4 * d
    return words, vecs


def write_word2vec_format(path: str, index2word, vectors: np.ndarray, binary: bool = True):
    vectors = np.ascontiguousarray(vectors, np.float32)
    with open(path, "wb") as f:
        f.write(f"{vectors.shape[0]} {vectors.shape[1]}\n".encode())
        for w, v in zip(index2word, vectors):
            if binary:
                f.write(w.encode("utf-8") + b" " + struct.pack(f"<{len(v)}f", *v) + b"\n")
            else:
                f.write((w + " " + " ".join(repr(float(x)) for x in v) + "\n").encode("utf-8"))
