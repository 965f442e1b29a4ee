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
