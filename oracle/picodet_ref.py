"""numpy restatement of the reference PicoDet post-processing (TEST ORACLE, see oracle/__init__.py).

Follows OCRPicodetPostProcessor.__call__ picodet/processor_picodet.py:184-298 (centres :207-214, DFL softmax-integral
:216-221, per-level top-k :223-228, decode :231, per-class threshold + NMS :240-256, warp_boxes / scale :262-272),
hard_nms :301-331, iou_of :334-351, area_of :354-360, warp_boxes :136-158 -- with the reference's dtypes: class scores
and the softmax stay float32, everything downstream of `softmax * arange` is float64, boxes pass through float32 once in
warp_boxes, and the clip uses the ORIGINAL image size before the division by the scale factor (bug-compatible).
Pinned against the reference class itself by tests/golden/picodet_post.npz (oracle/gen_golden_picodet.py).
"""
from __future__ import annotations

from typing import List, Sequence

import numpy as np


def softmax_f32(x: np.ndarray) -> np.ndarray:
    """scipy.special.softmax(x, axis=1) on float32."""
    x = x.astype(np.float32)
    e = np.exp(x - np.amax(x, axis=1, keepdims=True))
    return e / np.sum(e, axis=1, keepdims=True)


def _area(lt, rb):
    hw = np.clip(rb - lt, 0.0, None)
    return hw[..., 0] * hw[..., 1]


def _iou(boxes0, box1, eps=1e-5):
    lt = np.maximum(boxes0[..., :2], box1[..., :2])
    rb = np.minimum(boxes0[..., 2:], box1[..., 2:])
    ov = _area(lt, rb)
    return ov / (_area(boxes0[..., :2], boxes0[..., 2:]) + _area(box1[..., :2], box1[..., 2:]) - ov + eps)


def hard_nms(box_scores: np.ndarray, iou_threshold: float, top_k: int = -1, candidate_size: int = 200) -> np.ndarray:
    scores, boxes = box_scores[:, -1], box_scores[:, :-1]
    picked = []
    indexes = np.argsort(scores)[-candidate_size:]
    while len(indexes) > 0:
        current = indexes[-1]
        picked.append(current)
        if 0 < top_k == len(picked) or len(indexes) == 1:
            break
        cur = boxes[current, :]
        indexes = indexes[:-1]
        indexes = indexes[_iou(boxes[indexes, :], cur[None]) <= iou_threshold]
    return box_scores[picked, :]


def picodet_decode(scores: Sequence[np.ndarray], raw_boxes: Sequence[np.ndarray], org_shape, scale_factor, target_shape,
                   strides=(8, 16, 32, 64), score_threshold: float = 0.5, nms_threshold: float = 0.5, nms_top_k: int = 1000,
                   keep_top_k: int = 100) -> List[np.ndarray]:
    """scores[l] fp32 [B,HW_l,C], raw_boxes[l] fp32 [B,HW_l,4*(reg_max+1)] -> per image float64 [n,6] rows
    (class id, score, x1, y1, x2, y2), in the reference's order (classes ascending, NMS pick order inside a class)."""
    batch = raw_boxes[0].shape[0]
    reg_max = int(raw_boxes[0].shape[-1] / 4 - 1)
    ori_shape = np.array(org_shape, dtype=np.float32).reshape(-1, 2)
    sf = np.array(scale_factor, dtype=np.float32).reshape(-1, 2)
    out = []
    for b in range(batch):
        dec, sel = [], []
        for stride, dist, score in zip(strides, raw_boxes, scores):
            dist, score = dist[b], score[b]
            fm_h, fm_w = target_shape[0] / stride, target_shape[1] / stride
            ww, hh = np.meshgrid(np.arange(fm_w), np.arange(fm_h))
            ct_row, ct_col = (hh.flatten() + 0.5) * stride, (ww.flatten() + 0.5) * stride
            center = np.stack((ct_col, ct_row, ct_col, ct_row), axis=1)
            d = softmax_f32(dist.reshape((-1, reg_max + 1))) * np.expand_dims(np.arange(reg_max + 1), 0)
            d = np.sum(d, axis=1).reshape((-1, 4)) * stride
            idx = np.argsort(score.max(axis=1))[::-1][:nms_top_k]
            dec.append(center[idx] + [-1, -1, 1, 1] * d[idx])
            sel.append(score[idx])
        bboxes, conf = np.concatenate(dec, 0), np.concatenate(sel, 0)
        rows, labels = [], []
        for c in range(conf.shape[1]):
            probs = conf[:, c]
            mask = probs > score_threshold
            if not mask.any():
                continue
            bp = hard_nms(np.concatenate([bboxes[mask], probs[mask].reshape(-1, 1)], 1), nms_threshold, keep_top_k)
            rows.append(bp)
            labels += [c] * len(bp)
        if not rows:
            out.append(np.empty((0, 6)))
            continue
        bp = np.concatenate(rows)
        width, height = ori_shape[b][1], ori_shape[b][0]
        xy = bp[:, :4].copy()
        xy[:, [0, 2]] = xy[:, [0, 2]].clip(0, width)
        xy[:, [1, 3]] = xy[:, [1, 3]].clip(0, height)
        bp[:, :4] = xy.astype(np.float32)
        bp[:, :4] /= np.concatenate([sf[b][::-1], sf[b][::-1]])
        out.append(np.concatenate([np.array(labels, dtype=np.float64)[:, None], bp[:, 4:5], bp[:, :4]], 1))
    return out
