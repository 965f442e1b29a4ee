"""CPU restatement of the reference CenterNet table-structure path (TEST ORACLE, see oracle/__init__.py).

Network: center_net/modeling_centernet.py DLASeg.forward :657-662 = dla34 base (shared with Lore, oracle/lore_net_ref.py
dla34_base) -> DLAUp.forward :589-597 over plain IDAUp blocks :508-570 (proj 1x1+BN+ReLU unless channels match, depthwise
ConvTranspose2d up-sampling unless the factor is 1, node = 3x3 conv on cat[x, layer] + BN + ReLU) -> heads hm / v2c / c2v / reg.
Decode: OCRTableCenterNetPostProcessor.__call__ center_net/processer_centernet.py:170-204 with center_net/table_process.py
bbox_decode :151-185, gbox_decode :188-216, _nms / _topk :115-140, nms :239-275 (a no-op here: it is handed the [1,K,10]
batch array, so len(dets) < 2), bbox_post_process / gbox_post_process :219-236 (transform_preds = Lore's), group_bbox_by_gbox
:278-333, the score > 0.3 filter and the sort by 0.01 * mean_x + mean_y.
Deliberate restatement choices: `hm` is taken after the sigmoid; torch.topk tie order -> ascending index; rows below the 0.3
gates (which neither the grouping loops nor the final filter ever use) are not produced.  group_bbox_by_gbox is restated as the
equivalent "first (vertex, centre) in loop order wins each cell corner" rule and pinned against the reference's loops.
Pinned by tests/golden/centernet_*.npz (oracle/gen_golden_centernet.py).
"""
from __future__ import annotations

import math
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

from . import lore_decode_ref as L
from .lore_net_ref import _bn, _t, dla34_base

F32 = np.float32


@torch.no_grad()
def centernet_dla34_forward(sd, x: torch.Tensor) -> Dict[str, torch.Tensor]:
    layers = dla34_base(sd, x.float())[2:]
    channels = [64, 128, 256, 512]
    in_channels = list(channels)
    scales = np.array([1, 2, 4, 8], dtype=int)
    feat = None
    for i in range(3):
        j = -i - 2
        o, up_f = channels[j], scales[j:] // scales[j]
        p = f"dla_up.ida_{i}"
        ls = list(layers[j:])
        for k, l in enumerate(ls):
            if (f"{p}.proj_{k}.0.weight") in sd:
                l = F.relu(_bn(F.conv2d(l, _t(sd, f"{p}.proj_{k}.0.weight")), sd, f"{p}.proj_{k}.1"))
            f = int(up_f[k])
            if f != 1:
                w = _t(sd, f"{p}.up_{k}.weight")
                l = F.conv_transpose2d(l, w, stride=f, padding=f // 2, groups=w.shape[0])
            ls[k] = l
        xx, y = ls[0], []
        for k in range(1, len(ls)):
            xx = F.relu(_bn(F.conv2d(torch.cat([xx, ls[k]], 1), _t(sd, f"{p}.node_{k}.0.weight"), padding=1), sd, f"{p}.node_{k}.1"))
            y.append(xx)
        layers[-i - 1:] = y
        feat = xx
        scales[j + 1:] = scales[j]
        in_channels[j + 1:] = [channels[j]] * len(in_channels[j + 1:])
    out = {"feat": feat}
    for head in ("hm", "v2c", "c2v", "reg"):
        t = F.relu(F.conv2d(feat, _t(sd, f"{head}.0.weight"), _t(sd, f"{head}.0.bias"), padding=1))
        out[head] = F.conv2d(t, _t(sd, f"{head}.2.weight"), _t(sd, f"{head}.2.bias"))
    return out


def _dist(p, q) -> float:
    dx, dy = F32(p[0]) - F32(q[0]), F32(p[1]) - F32(q[1])
    return math.sqrt(F32(dx * dx + dy * dy))


def _point_in_box(b, px, py) -> bool:
    a = (b[2] - b[0]) * (py - b[1]) - (b[3] - b[1]) * (px - b[0])
    bb = (b[4] - b[2]) * (py - b[3]) - (b[5] - b[3]) * (px - b[2])
    c = (b[6] - b[4]) * (py - b[5]) - (b[7] - b[5]) * (px - b[4])
    d = (b[0] - b[6]) * (py - b[7]) - (b[1] - b[7]) * (px - b[6])
    return bool((a > 0 and bb > 0 and c > 0 and d > 0) or (a < 0 and bb < 0 and c < 0 and d < 0))


def group_cells_by_vertices(cells: np.ndarray, cell_scores: np.ndarray, verts: np.ndarray, vert_scores: np.ndarray, score_thred=0.3,
                            v2c_dist_thred=2.0, c2v_dist_thred=0.5) -> np.ndarray:
    """cells [n,8], verts [m,10] (vertex x, y, four predicted cell centres), both sorted by descending score, float32,
    source pixels.  Each corner of each cell snaps to the FIRST (vertex, centre) in loop order that claims it."""
    orig = cells.copy()
    out = cells.copy()
    sign = np.zeros((len(cells), 4), bool)
    for v, vs in zip(verts, vert_scores):
        if vs < F32(score_thred):
            break
        for i in range(4):
            cx, cy = v[2 * i + 2], v[2 * i + 3]
            if _dist(v[0:2], (cx, cy)) < v2c_dist_thred:
                continue
            for k in range(len(orig)):
                if cell_scores[k] < F32(score_thred):
                    break
                b = orig[k]
                if not _point_in_box(b, cx, cy):
                    continue
                w = (abs(b[6] - b[0]) + abs(b[4] - b[2])) / 2
                h = (abs(b[3] - b[1]) + abs(b[5] - b[7])) / 2
                m = max(w, h)
                dists = [_dist(v[0:2], (b[2 * j], b[2 * j + 1])) for j in range(4)]
                jm = int(np.argmin(dists))  # first minimum, as the reference's strict `<` scan
                if dists[jm] < 1e4 and dists[jm] < c2v_dist_thred * m and not sign[k, jm]:
                    out[k, 2 * jm], out[k, 2 * jm + 1] = v[0], v[1]
                    sign[k, jm] = True
    return out


def centernet_decode(hm, reg, c2v, v2c, center, scale, out_h, out_w, K=1000, MK=4000, score_thred=0.3) -> np.ndarray:
    """hm [2,H,W] AFTER sigmoid, reg [2,H,W], c2v / v2c [8,H,W] -> polygons float32 [n,8] in source pixels, in the reference's
    final order (sorted by 0.01 * mean_x + mean_y)."""
    hm, reg, c2v, v2c = (np.asarray(a, F32) for a in (hm, reg, c2v, v2c))
    H, W = hm.shape[1:]

    def peaks(heat, feat, k):
        s = L.nms_peaks(heat).reshape(-1)
        idx = np.flatnonzero(s >= F32(score_thred))
        order = np.lexsort((idx, -s[idx].astype(np.float64)))
        idx = idx[order][:k]
        xs = (idx % W).astype(F32) + reg.reshape(2, -1)[0, idx]
        ys = (idx // W).astype(F32) + reg.reshape(2, -1)[1, idx]
        g = feat.reshape(8, -1)[:, idx].T
        box = np.empty((len(idx), 8), F32)
        box[:, 0::2] = xs[:, None] - g[:, 0::2]
        box[:, 1::2] = ys[:, None] - g[:, 1::2]
        return s[idx].astype(F32), xs, ys, box

    cs, _, _, cbox = peaks(hm[0], c2v, K)
    vs, vx, vy, vbox = peaks(hm[1], v2c, MK)
    trans = L.affine_matrix(np.asarray(center, F32), F32(scale), int(out_w), int(out_h), True)
    cells = np.concatenate([L.transform_points(cbox[:, 2 * k:2 * k + 2], trans) for k in range(4)], 1)
    verts = np.concatenate([L.transform_points(np.stack([vx, vy], 1), trans)] + [L.transform_points(vbox[:, 2 * k:2 * k + 2], trans) for k in range(4)], 1)
    grouped = group_cells_by_vertices(cells, cs, verts, vs, score_thred)
    keep = grouped[cs > F32(score_thred)]
    key = [F32(0.01) * (F32(sum(b[::2])) / F32(4)) + F32(sum(b[1::2])) / F32(4) for b in keep]
    order = sorted(range(len(keep)), key=lambda i: key[i])
    return keep[order] if len(keep) else np.zeros((0, 8), F32)
