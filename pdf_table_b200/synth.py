"""Deterministic synthetic weights and inputs shared by the oracle, the tests and bench.py.

Weights are drawn from numpy's PCG64 with a fixed seed, in a fixed key order, so the build container
(where the reference is importable) and the GPU box (where it is not) see identical tensors without
shipping checkpoints.  Key names / shapes follow the reference modules' state_dicts.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np


def _conv(rng, cout, cin, kh, kw, gain=2.0):
    std = np.sqrt(gain / (cin * kh * kw))
    return (rng.standard_normal((cout, cin, kh, kw)) * std).astype(np.float32)


def _bn(rng, sd, prefix, c):
    sd[prefix + ".weight"] = rng.uniform(0.5, 1.5, c).astype(np.float32)
    sd[prefix + ".bias"] = (rng.standard_normal(c) * 0.1).astype(np.float32)
    sd[prefix + ".running_mean"] = (rng.standard_normal(c) * 0.1).astype(np.float32)
    sd[prefix + ".running_var"] = rng.uniform(0.5, 1.5, c).astype(np.float32)


def dbnet_r18_state_dict(seed: int = 0) -> "OrderedDict[str, np.ndarray]":
    """Keys of DBModel (reference model/db_net/dbnet.py:715-728) used in eval mode
    (the adaptive `thresh` branch :542-546 is never executed in eval and is omitted)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    sd["backbone.conv1.weight"] = _conv(rng, 64, 3, 7, 7)
    _bn(rng, sd, "backbone.bn1", 64)
    inpl = 64
    for L, planes in enumerate((64, 128, 256, 512), start=1):
        for B in range(2):
            p = f"backbone.layer{L}.{B}"
            stride = 2 if (L > 1 and B == 0) else 1
            sd[p + ".conv1.weight"] = _conv(rng, planes, inpl, 3, 3)
            _bn(rng, sd, p + ".bn1", planes)
            sd[p + ".conv2.weight"] = _conv(rng, planes, planes, 3, 3, gain=1.0)
            _bn(rng, sd, p + ".bn2", planes)
            if stride != 1 or inpl != planes:
                sd[p + ".downsample.0.weight"] = _conv(rng, planes, inpl, 1, 1, gain=1.0)
                _bn(rng, sd, p + ".downsample.1", planes)
            inpl = planes
    for name, cin in (("in5", 512), ("in4", 256), ("in3", 128), ("in2", 64)):
        sd[f"decoder.{name}.weight"] = _conv(rng, 256, cin, 1, 1, gain=1.0)
    for name in ("out5.0", "out4.0", "out3.0", "out2"):
        sd[f"decoder.{name}.weight"] = _conv(rng, 64, 256, 3, 3, gain=1.0)
    sd["decoder.binarize.0.weight"] = _conv(rng, 64, 256, 3, 3)
    _bn(rng, sd, "decoder.binarize.1", 64)
    sd["decoder.binarize.3.weight"] = (rng.standard_normal((64, 64, 2, 2)) * np.sqrt(2.0 / 64)).astype(np.float32)
    sd["decoder.binarize.3.bias"] = (rng.standard_normal(64) * 0.1).astype(np.float32)
    _bn(rng, sd, "decoder.binarize.4", 64)
    sd["decoder.binarize.6.weight"] = (rng.standard_normal((64, 1, 2, 2)) * (0.25 * np.sqrt(1.0 / 64))).astype(np.float32)
    sd["decoder.binarize.6.bias"] = (rng.standard_normal(1) * 0.1).astype(np.float32)
    return sd


def synthetic_page(index: int, h: int = 960, w: int = 960) -> np.ndarray:
    """white page, rendered text lines, optional ruled table, light noise.  uint8 HWC."""
    import cv2

    rng = np.random.default_rng(20240905 + index)
    img = np.full((h, w, 3), 255, np.uint8)
    chars = list("ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789 .,%-")
    n_lines = int(rng.integers(30, 61))
    y = 20
    for _ in range(n_lines):
        height = int(rng.integers(14, 29))
        scale = height / 22.0
        s = "".join(rng.choice(chars, size=int(rng.integers(8, 48))))
        x = int(rng.integers(10, max(11, w // 4)))
        col = int(rng.integers(0, 81))
        y += height + int(rng.integers(4, 12))
        if y >= h - 10:
            break
        cv2.putText(img, s, (x, y), cv2.FONT_HERSHEY_SIMPLEX, scale, (col, col, col), max(1, int(scale * 1.5)), cv2.LINE_AA)
    for _ in range(int(rng.integers(0, 3))):
        rows, cols = int(rng.integers(3, 9)), int(rng.integers(2, 7))
        tw, th = int(w * rng.uniform(0.5, 0.9)), int(h * rng.uniform(0.25, 0.5) * 0.5)
        x0, y0 = int(rng.integers(5, w - tw - 5)), int(rng.integers(5, h - th - 5))
        cv2.rectangle(img, (x0, y0), (x0 + tw, y0 + th), (255, 255, 255), -1)
        for r in range(rows + 1):
            yy = y0 + r * th // rows
            cv2.line(img, (x0, yy), (x0 + tw, yy), (0, 0, 0), int(rng.integers(1, 3)))
        for c in range(cols + 1):
            xx = x0 + c * tw // cols
            cv2.line(img, (xx, y0), (xx, y0 + th), (0, 0, 0), int(rng.integers(1, 3)))
    noise = rng.normal(0, 2, img.shape)
    return np.clip(img.astype(np.float32) + noise, 0, 255).astype(np.uint8)


def synthetic_text_crop(index: int, h: int = 32, w: int = 320) -> np.ndarray:
    """One rendered string on a light background, uint8 HWC (rec crop)."""
    import cv2

    rng = np.random.default_rng(20240905 + 100000 + index)
    img = np.full((h, w, 3), 255, np.uint8)
    chars = list("ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz0123456789 .,%-")
    s = "".join(rng.choice(chars, size=int(rng.integers(6, 26))))
    col = int(rng.integers(0, 81))
    cv2.putText(img, s, (4, int(h * 0.75)), cv2.FONT_HERSHEY_SIMPLEX, h / 40.0, (col, col, col), 1, cv2.LINE_AA)
    noise = rng.normal(0, 2, img.shape)
    return np.clip(img.astype(np.float32) + noise, 0, 255).astype(np.uint8)


# --------------------------------------------------------------------------- ConvNextViT
CONVNEXT_DEPTHS = (3, 3, 8, 3)
CONVNEXT_DIMS = (96, 192, 256, 512)
VIT_LAYERS, VIT_DIM, VIT_HEADS, VIT_MLP, VIT_TOKENS, VIT_LABELS = 12, 192, 3, 768, 75, 7644


def _lin(rng, cout, cin, gain=1.0):
    return (rng.standard_normal((cout, cin)) * np.sqrt(gain / cin)).astype(np.float32)


def _ln(rng, sd, prefix, c):
    sd[prefix + ".weight"] = rng.uniform(0.5, 1.5, c).astype(np.float32)
    sd[prefix + ".bias"] = (rng.standard_normal(c) * 0.1).astype(np.float32)


def _b(rng, c, s=0.1):
    return (rng.standard_normal(c) * s).astype(np.float32)


def convnext_vit_state_dict(seed: int = 0, num_labels: int = VIT_LABELS) -> "OrderedDict[str, np.ndarray]":
    """Keys / shapes of the reference ConvNextViT (model/convnext_vit/modeling_convnext_vit.py:20-45:
    ConvNeXt depths [3,3,8,3] dims [96,192,256,512] on 1 channel, (2,1) down-sampling; ViT 12 x 192 x 3 heads
    over 75 tokens; classifier 192 -> 7644).  layer_scale is drawn O(0.1..0.5) (its 1e-6 init would hide
    every block); unused parameters (cls_token, cnn_model.layernorm, position 0) are still drawn so the
    dict loads strictly into the reference module."""
    rng = np.random.Generator(np.random.PCG64(1000 + seed))
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    p = "cnn_model.embeddings"
    sd[p + ".patch_embeddings.weight"] = _conv(rng, 96, 1, 4, 4, gain=1.0)
    sd[p + ".patch_embeddings.bias"] = _b(rng, 96)
    _ln(rng, sd, p + ".layernorm", 96)
    prev = CONVNEXT_DIMS[0]
    for s, (depth, dim) in enumerate(zip(CONVNEXT_DEPTHS, CONVNEXT_DIMS)):
        sp = f"cnn_model.encoder.stages.{s}"
        if s > 0:
            _ln(rng, sd, sp + ".downsampling_layer.0", prev)
            sd[sp + ".downsampling_layer.1.weight"] = _conv(rng, dim, prev, 2, 1, gain=1.0)
            sd[sp + ".downsampling_layer.1.bias"] = _b(rng, dim)
        for j in range(depth):
            lp = f"{sp}.layers.{j}"
            sd[lp + ".layer_scale_parameter"] = rng.uniform(0.1, 0.5, dim).astype(np.float32)
            sd[lp + ".dwconv.weight"] = _conv(rng, dim, 1, 7, 7, gain=1.0)
            sd[lp + ".dwconv.bias"] = _b(rng, dim)
            _ln(rng, sd, lp + ".layernorm", dim)
            sd[lp + ".pwconv1.weight"] = _lin(rng, 4 * dim, dim, gain=2.0)
            sd[lp + ".pwconv1.bias"] = _b(rng, 4 * dim)
            sd[lp + ".pwconv2.weight"] = _lin(rng, dim, 4 * dim)
            sd[lp + ".pwconv2.bias"] = _b(rng, dim)
        prev = dim
    _ln(rng, sd, "cnn_model.layernorm", 512)
    v = "vitstr.vit"
    sd[v + ".embeddings.cls_token"] = _b(rng, VIT_DIM, 0.02).reshape(1, 1, VIT_DIM)
    sd[v + ".embeddings.position_embeddings"] = _b(rng, (VIT_TOKENS + 1) * VIT_DIM, 0.2).reshape(1, VIT_TOKENS + 1, VIT_DIM)
    sd[v + ".embeddings.patch_embeddings.projection.weight"] = _conv(rng, VIT_DIM, 512, 1, 1, gain=1.0)
    sd[v + ".embeddings.patch_embeddings.projection.bias"] = _b(rng, VIT_DIM)
    for L in range(VIT_LAYERS):
        lp = f"{v}.encoder.layer.{L}"
        for n in ("query", "key", "value"):
            sd[f"{lp}.attention.attention.{n}.weight"] = _lin(rng, VIT_DIM, VIT_DIM)
            sd[f"{lp}.attention.attention.{n}.bias"] = _b(rng, VIT_DIM)
        sd[lp + ".attention.output.dense.weight"] = _lin(rng, VIT_DIM, VIT_DIM, gain=0.5)
        sd[lp + ".attention.output.dense.bias"] = _b(rng, VIT_DIM)
        sd[lp + ".intermediate.dense.weight"] = _lin(rng, VIT_MLP, VIT_DIM, gain=2.0)
        sd[lp + ".intermediate.dense.bias"] = _b(rng, VIT_MLP)
        sd[lp + ".output.dense.weight"] = _lin(rng, VIT_DIM, VIT_MLP, gain=0.5)
        sd[lp + ".output.dense.bias"] = _b(rng, VIT_DIM)
        _ln(rng, sd, lp + ".layernorm_before", VIT_DIM)
        _ln(rng, sd, lp + ".layernorm_after", VIT_DIM)
    _ln(rng, sd, v + ".layernorm", VIT_DIM)
    sd["vitstr.classifier.weight"] = _lin(rng, num_labels, VIT_DIM, gain=4.0)
    sd["vitstr.classifier.bias"] = _b(rng, num_labels)
    return sd


# --------------------------------------------------------------------------- planted DB probability maps
def synthetic_prob_map(index: int, h: int = 960, w: int = 960, n_lines: int = 40) -> np.ndarray:
    """A DB-like probability map (fp32 [h,w] in [0,1)) with analytically placed text-line blobs: rotated soft-edged
    rectangles of varying peak probability (some below box_thresh), a few with holes, a ruled-table frame, and
    small specks.  numpy only, so the build container and the GPU box produce identical maps."""
    rng = np.random.default_rng(424242 + index)
    prob = np.zeros((h, w), np.float32)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)

    def add_rect(cx, cy, bw, bh, ang, peak, soft, hole=None):
        c, s = np.float32(np.cos(ang)), np.float32(np.sin(ang))
        r = int(np.hypot(bw, bh) / 2 + soft + 2)
        x0, x1, y0, y1 = max(0, int(cx) - r), min(w, int(cx) + r + 1), max(0, int(cy) - r), min(h, int(cy) + r + 1)
        if x0 >= x1 or y0 >= y1:
            return
        dx, dy = xx[y0:y1, x0:x1] - np.float32(cx), yy[y0:y1, x0:x1] - np.float32(cy)
        u, v = dx * c + dy * s, -dx * s + dy * c
        d = np.minimum(np.float32(bw / 2) - np.abs(u), np.float32(bh / 2) - np.abs(v))  # >0 inside
        val = np.float32(peak) * np.clip(np.float32(0.5) + d / np.float32(soft), 0, 1)
        if hole is not None:
            hu, hv, hr = hole
            val = np.where((u - np.float32(hu)) ** 2 + (v - np.float32(hv)) ** 2 < np.float32(hr * hr), np.float32(0.05), val)
        prob[y0:y1, x0:x1] = np.maximum(prob[y0:y1, x0:x1], val.astype(np.float32))

    placed = []  # (x0, y0, x1, y1) of accepted blobs incl. margin: blobs never merge

    def try_place(cx, cy, bw, bh, ang, margin=10.0):
        ex = abs(np.cos(ang)) * bw / 2 + abs(np.sin(ang)) * bh / 2 + margin
        ey = abs(np.sin(ang)) * bw / 2 + abs(np.cos(ang)) * bh / 2 + margin
        box = (cx - ex, cy - ey, cx + ex, cy + ey)
        if box[0] < -5 or box[1] < -5 or box[2] > w + 5 or box[3] > h + 5:
            return rng.random() < 0.1  # a few blobs may touch the frame
        for q in placed:
            if box[0] < q[2] and q[0] < box[2] and box[1] < q[3] and q[1] < box[3]:
                return False
        placed.append(box)
        return True

    # ruled table frame: thin lines forming cells (one big component with many holes)
    tx, ty, tw, th = rng.uniform(0.1, 0.3) * w, rng.uniform(0.55, 0.7) * h, rng.uniform(0.4, 0.6) * w, rng.uniform(0.15, 0.25) * h
    rows, cols = int(rng.integers(2, 5)), int(rng.integers(2, 6))
    for r in range(rows + 1):
        add_rect(tx + tw / 2, ty + r * th / rows, tw, 3.0, 0.0, 0.9, 1.0)
    for c in range(cols + 1):
        add_rect(tx + c * tw / cols, ty + th / 2, 3.0, th, 0.0, 0.9, 1.0)
    placed.append((tx - 12, ty - 12, tx + tw + 12, ty + th + 12))
    n_ok = 0
    for i in range(n_lines * 30):
        if n_ok >= n_lines:
            break
        bw, bh = rng.uniform(40, min(420, w * 0.45)), rng.uniform(8, 30)
        ang = rng.uniform(-0.2, 0.2) if rng.random() < 0.8 else rng.uniform(-1.5, 1.5)
        if rng.random() < 0.35:
            ang = 0.0
        cx, cy = rng.uniform(0, w), rng.uniform(0, h)
        peak = rng.uniform(0.5, 0.98)
        hole = (rng.uniform(-bw / 4, bw / 4), 0.0, rng.uniform(1.5, 4.0)) if rng.random() < 0.15 else None
        soft = rng.uniform(1.0, 3.0)
        if not try_place(cx, cy, bw, bh, ang):
            continue
        add_rect(cx, cy, bw, bh, ang, peak, soft, hole)
        n_ok += 1
    for i in range(25):  # specks: tiny components that fail min_size or box_thresh
        add_rect(rng.uniform(5, w - 5), rng.uniform(5, h - 5), rng.uniform(1, 5), rng.uniform(1, 5), 0.0, rng.uniform(0.3, 0.9), 1.0)
    return prob
