"""CPU restatement of the reference Lore detect-output decoding (TEST ORACLE, see oracle/__init__.py).

Follows lore/lineless_table_process.py: process_detect_output :592-655, corner_decode :97-124,
ctdet_4ps_decode :127-267 (incl. the sequential `wiz_rev` corner snapping :178-236), _nms :66-73, _topk :76-94,
_get_4ps_feat :39-63 (with its odd upper clamp `feat.shape[0]-1` = 0), is_group_faster_faster :355-379,
find4ps :329-338, dist :341-345, ctdet_4ps_post_process / transform_preds / affine_transform :489-507, 471-476,
387-390, get_affine_transform :403-438, merge_outputs :551-565, filter :568-582, normalized_ps :585-589, and
process_logic_output :658-663.  float32 arithmetic is kept float32 (numpy scalars), float64 where the reference
promotes (np.dot with the float64 cv2 affine matrix).

Deliberate restatement choices (documented, also in DESIGN.md):
  * `hm` is taken AFTER the sigmoid (the reference applies `sigmoid_()` first thing, :599); the engine fuses the
    sigmoid into the head epilogue.
  * torch.topk / torch.sort leave the order of exactly-equal scores unspecified; here ties are broken by ascending
    flat index (topk) and ascending pre-sort rank (sort).  Rows with score 0 (NMS-suppressed padding of the
    reference's K=3000 / MK=5000 lists) carry arbitrary positions in the reference and are never selected
    (vis_thresh > 0); only rows with a positive score are produced here.
  * shapely is not installed in the build image: `Point.within(Polygon)` is restated as a strict point-in-polygon
    test (crossing number, boundary excluded) in float64 -- parity UNPINNED for that third-party predicate.
Pinned against the reference functions themselves (run with `.cuda()` neutralised and the shapely stub above) by
tests/golden/lore_decode.npz (oracle/gen_golden_lore.py).
"""
from __future__ import annotations

from typing import Dict

import cv2
import numpy as np

F32 = np.float32


def nms_peaks(heat: np.ndarray) -> np.ndarray:
    """heat fp32 [H,W] -> heat * (maxpool3x3(heat) == heat), -inf padding as F.max_pool2d."""
    h, w = heat.shape
    p = np.full((h + 2, w + 2), -np.inf, F32)
    p[1:-1, 1:-1] = heat
    m = heat.copy()
    for dy in range(3):
        for dx in range(3):
            m = np.maximum(m, p[dy:dy + h, dx:dx + w])
    return np.where(m == heat, heat, F32(0))


def topk_peaks(heat: np.ndarray, k: int):
    """-> (scores desc, flat indices) of the (at most k) positive NMS survivors; ties by ascending index."""
    s = nms_peaks(heat).reshape(-1)
    idx = np.flatnonzero(s > 0)
    order = np.lexsort((idx, -s[idx].astype(np.float64)))
    idx = idx[order][:k]
    return s[idx].astype(F32), idx.astype(np.int64)


def _gather(feat_chw: np.ndarray, idx: np.ndarray) -> np.ndarray:
    c = feat_chw.shape[0]
    return feat_chw.reshape(c, -1)[:, idx].T.astype(F32)  # [n, c]


def point_strictly_in_polygon(px: float, py: float, poly: np.ndarray) -> bool:
    """Strict interior test (float64) for a simple polygon given as [n,2]; points on the boundary are outside."""
    n = len(poly)
    inside = False
    for a in range(n):
        x1, y1 = float(poly[a][0]), float(poly[a][1])
        x2, y2 = float(poly[(a + 1) % n][0]), float(poly[(a + 1) % n][1])
        # on-segment -> boundary -> not within
        cross = (x2 - x1) * (py - y1) - (y2 - y1) * (px - x1)
        if cross == 0.0 and min(x1, x2) <= px <= max(x1, x2) and min(y1, y2) <= py <= max(y1, y2):
            return False
        if (y1 > py) != (y2 > py):
            # x coordinate of the edge at height py, compared without division: sign-aware
            t = (px - x1) * (y2 - y1) - (x2 - x1) * (py - y1)
            if (t < 0) == ((y2 - y1) > 0):
                inside = not inside
    return inside


def is_group(bbox: np.ndarray, gbox: np.ndarray) -> bool:
    b = bbox.reshape(4, 2)
    g = gbox.reshape(4, 2)
    if b[:, 0].min() > g[:, 0].max() or g[:, 0].min() > b[:, 0].max() or b[:, 1].min() > g[:, 1].max() or \
            g[:, 1].min() > b[:, 1].max():
        return False
    for i in range(4):
        if point_strictly_in_polygon(g[i, 0], g[i, 1], b):
            return True
    return False


def find4ps(bbox: np.ndarray, x: F32, y: F32) -> int:
    dx = bbox[0::2].astype(F32) - F32(x)
    dy = bbox[1::2].astype(F32) - F32(y)
    return int(np.argmin(dx * dx + dy * dy))


def _dist(x1, y1, x2, y2) -> F32:
    dx = F32(x1) - F32(x2)
    dy = F32(y1) - F32(y2)
    return F32(dx * dx + dy * dy)


def affine_matrix(center, scale, out_w, out_h, inv: bool) -> np.ndarray:
    """get_affine_transform(center, scale, 0, (out_w, out_h), inv) :403-438 (rot = 0, shift = 0)."""
    c = np.asarray(center, F32)
    src = np.zeros((3, 2), F32)
    dst = np.zeros((3, 2), F32)
    src_w = F32(scale)
    src[0] = c
    src[1] = c + np.array([0.0 * 1.0 - (src_w * -0.5) * 0.0, 0.0 * 0.0 + (src_w * -0.5) * 1.0], F32)
    dst[0] = [out_w * 0.5, out_h * 0.5]
    dst[1] = np.array([out_w * 0.5, out_h * 0.5], F32) + np.array([0, out_w * -0.5], F32)
    d = src[0] - src[1]
    src[2] = src[1] + np.array([-d[1], d[0]], F32)
    d = dst[0] - dst[1]
    dst[2] = dst[1] + np.array([-d[1], d[0]], F32)
    return cv2.getAffineTransform(dst, src) if inv else cv2.getAffineTransform(src, dst)


def upper_left_matrix(center, scale, out_w, out_h, inv: bool) -> np.ndarray:
    """get_affine_transform_upper_left :441-468."""
    src = np.zeros((3, 2), F32)
    dst = np.zeros((3, 2), F32)
    src[0] = center
    if center[0] < center[1]:
        src[1] = [scale, center[1]]
        dst[1] = [out_w, 0]
    else:
        src[1] = [center[0], scale]
        dst[1] = [0, out_w]
    d = src[0] - src[1]
    src[2] = src[1] + np.array([-d[1], d[0]], F32)
    d = dst[0] - dst[1]
    dst[2] = dst[1] + np.array([-d[1], d[0]], F32)
    return cv2.getAffineTransform(dst, src) if inv else cv2.getAffineTransform(src, dst)


def transform_points(pts: np.ndarray, trans: np.ndarray) -> np.ndarray:
    """affine_transform :387-390 per point: float64 dot of the float64 matrix with [x, y, 1] (float32 inputs),
    result stored back into a float32 array."""
    p = np.concatenate([pts.astype(F32), np.ones((len(pts), 1), F32)], 1).astype(np.float64)
    return (p @ trans.T.astype(np.float64)).astype(F32)


def lore_decode(hm: np.ndarray, reg: np.ndarray, wh: np.ndarray, st: np.ndarray, ax: np.ndarray, cr: np.ndarray,
                meta, upper_left: bool = False, wiz_rev: bool = True, vis_thresh: float = 0.2, K: int = 3000,
                MK: int = 5000, batch_clamp: int = 0) -> Dict[str, np.ndarray]:
    """One image.  hm [2,H,W] AFTER sigmoid, reg [2,H,W], wh [8,H,W], st [8,H,W], ax / cr [256,H,W] (fp32);
    meta = the reference's int64 [cx, cy, s, in_h, in_w, out_h, out_w].
    -> logi_feat [n,256], dets_feat int64 [n,8], polygons [n,8] (source pixels), scores [n], plus the full
    positive-score lists for inspection."""
    hm, reg, wh, st = (np.asarray(a, F32) for a in (hm, reg, wh, st))
    H, W = hm.shape[1:]
    # ---- corner_decode on class 1
    c_scores, c_inds = topk_peaks(hm[1], MK)
    c_xs = (c_inds % W).astype(F32) + _gather(reg, c_inds)[:, 0]
    c_ys = (c_inds // W).astype(F32) + _gather(reg, c_inds)[:, 1]
    st_g = _gather(st, c_inds)
    gboxes = np.empty((len(c_inds), 8), F32)
    gboxes[:, 0::2] = c_xs[:, None] - st_g[:, 0::2]
    gboxes[:, 1::2] = c_ys[:, None] - st_g[:, 1::2]
    # ---- cells on class 0
    scores, inds = topk_peaks(hm[0], K)
    scores = scores.copy()
    xs = (inds % W).astype(F32) + _gather(reg, inds)[:, 0]
    ys = (inds // W).astype(F32) + _gather(reg, inds)[:, 1]
    wh_g = _gather(wh, inds)
    bboxes = np.empty((len(inds), 8), F32)
    bboxes[:, 0::2] = xs[:, None] - wh_g[:, 0::2]
    bboxes[:, 1::2] = ys[:, None] - wh_g[:, 1::2]
    rev = bboxes.copy()
    if wiz_rev:
        for i in range(len(inds)):
            if not scores[i] >= F32(0.2):
                break
            count = 0
            for j in range(len(c_inds)):
                if not c_scores[j] >= F32(0.3):
                    break
                if not is_group(bboxes[i], gboxes[j]):
                    continue
                cx, cy = c_xs[j], c_ys[j]
                q = find4ps(bboxes[i], cx, cy)
                if rev[i, 2 * q] == bboxes[i, 2 * q] and rev[i, 2 * q + 1] == bboxes[i, 2 * q + 1]:
                    count += 1
                    rev[i, 2 * q], rev[i, 2 * q + 1] = cx, cy
                elif _dist(bboxes[i, 2 * q], bboxes[i, 2 * q + 1], rev[i, 2 * q], rev[i, 2 * q + 1]) >= \
                        _dist(bboxes[i, 2 * q], bboxes[i, 2 * q + 1], cx, cy):
                    count += 1
                    rev[i, 2 * q], rev[i, 2 * q + 1] = cx, cy
            if count <= 2:
                scores[i] = F32(scores[i] * F32(0.4))
    # ---- cc_match (float32 arithmetic, round-half-even) and the corner-feature gather
    cc = np.rint(rev[:, 0::2] + F32(W) * np.rint(rev[:, 1::2])).astype(np.int64)
    cc = np.where(cc < H * W, cc, batch_clamp)
    cc = np.where(cc >= 0, cc, 0)
    cr_flat = np.asarray(cr, F32).reshape(cr.shape[0], -1)
    cr_feat = np.zeros((len(inds), cr.shape[0]), F32)
    for c in range(4):
        cr_feat = cr_feat + cr_flat[:, cc[:, c]].T
    ax_g = _gather(np.asarray(ax, F32), inds)
    if wiz_rev:
        order = np.lexsort((np.arange(len(inds)), -scores.astype(np.float64)))
        dets = rev[order]
        sorted_scores = scores[order]
        ax_g = ax_g[order]
    else:
        order = np.arange(len(inds))
        dets = bboxes
        sorted_scores = scores
    logi = ax_g + cr_feat  # reference quirk: ax is re-sorted, cr_feat is NOT (:255-262, :644)
    # ---- to source pixels
    c = [F32(meta[0]), F32(meta[1])]
    s = F32(meta[2])
    out_h, out_w = int(meta[5]), int(meta[6])
    trans = upper_left_matrix(c, s, out_w, out_h, True) if upper_left else affine_matrix(c, s, out_w, out_h, True)
    poly = np.empty_like(dets)
    for k in range(4):
        poly[:, 2 * k:2 * k + 2] = transform_points(dets[:, 2 * k:2 * k + 2], trans)
    n = int((sorted_scores >= F32(vis_thresh)).sum())
    dets_feat = np.clip(np.trunc(dets[:n]).astype(np.int64), 0, 255)  # int32 truncation in filter(), then normalized_ps
    return {
        "logi_feat": logi[:n], "dets_feat": dets_feat, "polygons": poly[:n], "scores": sorted_scores[:n],
        "all_polygons": poly, "all_scores": sorted_scores, "order": order, "cc_match": cc, "rev": rev,
        "bboxes": bboxes, "cell_inds": inds, "corner_inds": c_inds, "trans": trans,
    }


def round_logic(logi: np.ndarray) -> np.ndarray:
    """process_logic_output :658-663: floor + 1 if frac > 0.5 else floor (round-half-DOWN)."""
    f = np.floor(logi)
    return np.where(logi - f > 0.5, f + 1, f)
