"""numpy restatement of the reference CTC greedy decode (TEST ORACLE, see oracle/__init__.py).

Follows CTCLabelDecode.__call__ (model/ocr_rec_pp/rec_postprocess.py:175-191) and
BaseRecLabelDecode.decode(is_remove_duplicate=True) (:126-161); blank = 0 (:163-165).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import numpy as np


def ctc_greedy_ids(preds: np.ndarray, blank: int = 0) -> Tuple[List[np.ndarray], np.ndarray]:
    """preds [B,T,C] float32 -> (list of kept id arrays, conf[B] float32)."""
    preds = np.asarray(preds)
    idx = preds.argmax(axis=2)          # rec_postprocess.py:180
    prob = preds.max(axis=2)            # :181
    ids, conf = [], np.zeros(preds.shape[0], np.float32)
    for b in range(preds.shape[0]):
        sel = np.ones(idx.shape[1], dtype=bool)
        sel[1:] = idx[b][1:] != idx[b][:-1]     # :134-136
        sel &= idx[b] != blank                  # :137-138
        ids.append(idx[b][sel].astype(np.int32))
        c = prob[b][sel]
        if len(c) == 0:                         # :148-149
            c = [0]
        conf[b] = np.mean(c)                    # :157  (float32 pairwise sum / n)
    return ids, conf


def ctc_decode_text(preds: np.ndarray, character: Sequence[str]) -> List[Tuple[str, float]]:
    """Full (text, confidence) result; `character` is the dictionary with 'blank' prepended (:193-195)."""
    ids, conf = ctc_greedy_ids(preds)
    return [("".join(character[i] for i in row), float(c)) for row, c in zip(ids, conf)]
