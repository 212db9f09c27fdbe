"""Style-transfer-intensity metric on the B200 engine: the drop-in for the reference's
evaluate/auto/transfer_intensity.py, whose only arithmetic is a direct ``pyemd.emd`` call on two
class-probability vectors with an all-ones ground matrix (transfer_intensity.py:8-11).

Same function names, arguments and return values; ``calculate_STIs`` scores the whole list with ONE
batched call of ``wmd_emd_batch_host`` instead of one pyemd call per sentence.  There is no CPU path:
without a CUDA device these raise ``RuntimeError``.
"""
from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

from .engine import WMDEngine

_engine: Optional[WMDEngine] = None


def _eng(engine: Optional[WMDEngine] = None) -> WMDEngine:
    """The emd entry needs a handle (device, stream, workspace) but no embedding table."""
    global _engine
    if engine is not None:
        return engine
    if _engine is None:
        _engine = WMDEngine(np.ones((1, 1), np.float32))
    return _engine


def calculate_emd(input_distribution, output_distribution, engine: Optional[WMDEngine] = None) -> float:
    # transfer_intensity.py:8-11
    p = np.asarray(input_distribution, np.float64).reshape(1, -1)
    q = np.asarray(output_distribution, np.float64).reshape(1, -1)
    N = p.shape[1]
    return float(_eng(engine).emd_batch(p, q, np.ones((N, N)))[0])


def account_for_direction(input_target_style_probability, output_target_style_probability) -> int:
    # transfer_intensity.py:13-16
    if output_target_style_probability >= input_target_style_probability:
        return 1
    return -1


def calculate_direction_corrected_emd(input_distribution, output_distribution, target_style_class,
                                      engine: Optional[WMDEngine] = None) -> float:
    # transfer_intensity.py:18-21
    emd_score = calculate_emd(input_distribution, output_distribution, engine)
    return emd_score * account_for_direction(input_distribution[target_style_class], output_distribution[target_style_class])


def direction_corrected_emds(input_probs, output_probs, target_styles, engine: Optional[WMDEngine] = None):
    """Batched form: [B, N] probability matrices -> list of B signed scores."""
    P = np.asarray(input_probs, np.float64); Q = np.asarray(output_probs, np.float64)
    if P.ndim != 2 or P.shape != Q.shape:
        raise ValueError("expected two [B, N] arrays")
    if P.shape[0] == 0:
        return []
    N = P.shape[1]
    emd = _eng(engine).emd_batch(P, Q, np.ones((N, N)))
    tgt = np.asarray(target_styles, np.int64)
    rows = np.arange(P.shape[0])
    sign = np.where(Q[rows, tgt] >= P[rows, tgt], 1.0, -1.0)
    return [float(v) for v in emd * sign]


def calculate_STIs(sequences_input: Sequence[str], sequences_output: Sequence[str], target_styles, model,
                   engine: Optional[WMDEngine] = None):
    # transfer_intensity.py:23-32; `model` is a fasttext classifier (predict(sequence, k) -> labels, probabilities)
    def get_class_probs(sequence, model):
        labels, ps = model.predict(sequence, k=len(model.labels))
        pairs = list(zip(labels, np.asarray(ps).tolist()))
        pairs.sort(key=lambda e: e[0])
        return np.array([p for _, p in pairs])
    input_probs = [get_class_probs(s, model) for s in sequences_input]
    output_probs = [get_class_probs(s, model) for s in sequences_output]
    return direction_corrected_emds(input_probs, output_probs, target_styles, engine)
