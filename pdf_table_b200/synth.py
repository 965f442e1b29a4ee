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


def synthetic_prob_map_lines(index: int, h: int = 960, w: int = 960, n_lines: int = 40) -> np.ndarray:
    """A DB-like probability map whose every blob becomes a box: n_lines soft-edged text-line rectangles on a jittered
    column / row grid (slightly rotated, peak 0.8-0.98, none touching), plus specks that the size / score filters reject.
    The bench plants it so that the number of detected boxes per page is the workload's crop count."""
    rng = np.random.default_rng(515151 + index)
    prob = np.zeros((h, w), np.float32)
    yy, xx = np.mgrid[0:h, 0:w].astype(np.float32)
    cols = 4
    rows = (n_lines + cols - 1) // cols
    cw, rh = w / cols, h / rows
    for k in range(n_lines):
        r, c = divmod(k, cols)
        bw, bh = rng.uniform(0.55, 0.8) * cw, rng.uniform(0.28, 0.42) * rh
        cx = (c + 0.5) * cw + rng.uniform(-0.06, 0.06) * cw
        cy = (r + 0.5) * rh + rng.uniform(-0.1, 0.1) * rh
        ang = rng.uniform(-0.035, 0.035)
        peak, soft = rng.uniform(0.8, 0.98), rng.uniform(1.0, 2.5)
        cs, sn = np.float32(np.cos(ang)), np.float32(np.sin(ang))
        rad = int(np.hypot(bw, bh) / 2 + soft + 2)
        x0, x1, y0, y1 = max(0, int(cx) - rad), min(w, int(cx) + rad + 1), max(0, int(cy) - rad), min(h, int(cy) + rad + 1)
        dx, dy = xx[y0:y1, x0:x1] - np.float32(cx), yy[y0:y1, x0:x1] - np.float32(cy)
        u, v = dx * cs + dy * sn, -dx * sn + dy * cs
        d = np.minimum(np.float32(bw / 2) - np.abs(u), np.float32(bh / 2) - np.abs(v))
        val = np.float32(peak) * np.clip(np.float32(0.5) + d / np.float32(soft), 0, 1)
        prob[y0:y1, x0:x1] = np.maximum(prob[y0:y1, x0:x1], val.astype(np.float32))
    for _ in range(20):  # specks between the lines: tiny components that fail min_size
        sx, sy = int(rng.uniform(2, w - 4)), int(rng.uniform(2, h - 4))
        if prob[max(0, sy - 6):sy + 8, max(0, sx - 6):sx + 8].max() == 0:
            prob[sy:sy + 2, sx:sx + 2] = np.float32(rng.uniform(0.3, 0.9))
    return prob


# --------------------------------------------------------------------------- Lore (DLA-34 + DCNv2 detector, wtw)
DLA_LEVELS = (1, 1, 1, 2, 2, 1)
DLA_CHANNELS = (16, 32, 64, 128, 256, 512)
LORE_HEADS = (("hm", 2), ("st", 8), ("wh", 8), ("ax", 256), ("cr", 256), ("reg", 2))


def _dla_block(rng, sd, p, cin, cout):
    sd[p + ".conv1.weight"] = _conv(rng, cout, cin, 3, 3)
    _bn(rng, sd, p + ".bn1", cout)
    sd[p + ".conv2.weight"] = _conv(rng, cout, cout, 3, 3, gain=1.0)
    _bn(rng, sd, p + ".bn2", cout)


def _dla_tree(rng, sd, p, levels, cin, cout, stride, level_root=False, root_dim=0):
    """Parameter order of the reference Tree (center_net/modeling_centernet.py:209-287): tree1, tree2, root, project."""
    if root_dim == 0:
        root_dim = 2 * cout
    if level_root:
        root_dim += cin
    if levels == 1:
        _dla_block(rng, sd, p + ".tree1", cin, cout)
        _dla_block(rng, sd, p + ".tree2", cout, cout)
        sd[p + ".root.conv.weight"] = _conv(rng, cout, root_dim, 1, 1)
        _bn(rng, sd, p + ".root.bn", cout)
    else:
        _dla_tree(rng, sd, p + ".tree1", levels - 1, cin, cout, stride)
        _dla_tree(rng, sd, p + ".tree2", levels - 1, cout, cout, 1, root_dim=root_dim + cout)
    if cin != cout:
        sd[p + ".project.0.weight"] = _conv(rng, cout, cin, 1, 1, gain=1.0)
        _bn(rng, sd, p + ".project.1", cout)


def _dcn(rng, sd, p, cin, cout, off_std=0.04):
    """DeformConv (lore_dla_34.py:65-85) = DCN (dcnv2.py:25-86) + BN + ReLU.  The reference zero-initialises
    conv_offset_mask; here it is drawn so that the sampling path is exercised with offsets of the size a trained
    DCN produces (about one pixel: a per-tap bias of std 0.7 px plus a small data-dependent part) -- random O(1)
    weights on these un-normalised synthetic activations would give offsets of tens of pixels, i.e. a network whose
    output is chaotic in its own rounding noise."""
    _bn(rng, sd, p + ".actf.0", cout)
    sd[p + ".conv.weight"] = _conv(rng, cout, cin, 3, 3)
    sd[p + ".conv.bias"] = _b(rng, cout)
    sd[p + ".conv.conv_offset_mask.weight"] = (_conv(rng, 27, cin, 3, 3, gain=1.0) * off_std).astype(np.float32)
    sd[p + ".conv.conv_offset_mask.bias"] = np.concatenate([_b(rng, 18, 0.7), _b(rng, 9, 0.5)])


def _bilinear_up(c, f):
    """fill_up_weights (lore_dla_34.py:51-62): the fixed bilinear kernel every depthwise ConvTranspose2d starts from."""
    k = 2 * f
    ff = int(np.ceil(k / 2))
    cc = (2 * ff - 1 - ff % 2) / (2.0 * ff)
    w1 = np.array([1 - abs(i / ff - cc) for i in range(k)], np.float64)
    w = np.outer(w1, w1).astype(np.float32)
    return np.broadcast_to(w, (c, 1, k, k)).copy()


def _ida(rng, sd, p, o, channels, up_f, perturb_up):
    for i in range(1, len(channels)):
        _dcn(rng, sd, f"{p}.proj_{i}", channels[i], o)
        w = _bilinear_up(o, int(up_f[i]))
        if perturb_up:  # trained checkpoints may carry non-bilinear kernels: keep the kernel general
            w = (w * rng.uniform(0.8, 1.2, w.shape)).astype(np.float32)
        sd[f"{p}.up_{i}.weight"] = w
        _dcn(rng, sd, f"{p}.node_{i}", o, o)


def lore_dla34_state_dict(seed: int = 0, perturb_up: bool = True) -> "OrderedDict[str, np.ndarray]":
    """Keys / shapes of `get_dla_dcn(34, heads, head_conv=256)` (reference model/lore/lore_dla_34.py:140-206 over
    `dla34` center_net/modeling_centernet.py:289-402).  `base.fc` (unused in the forward) is omitted."""
    rng = np.random.Generator(np.random.PCG64(2000 + seed))
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    ch = DLA_CHANNELS
    sd["base.base_layer.0.weight"] = _conv(rng, ch[0], 3, 7, 7)
    _bn(rng, sd, "base.base_layer.1", ch[0])
    sd["base.level0.0.weight"] = _conv(rng, ch[0], ch[0], 3, 3)
    _bn(rng, sd, "base.level0.1", ch[0])
    sd["base.level1.0.weight"] = _conv(rng, ch[1], ch[0], 3, 3)
    _bn(rng, sd, "base.level1.1", ch[1])
    _dla_tree(rng, sd, "base.level2", DLA_LEVELS[2], ch[1], ch[2], 2, level_root=False)
    _dla_tree(rng, sd, "base.level3", DLA_LEVELS[3], ch[2], ch[3], 2, level_root=True)
    _dla_tree(rng, sd, "base.level4", DLA_LEVELS[4], ch[3], ch[4], 2, level_root=True)
    _dla_tree(rng, sd, "base.level5", DLA_LEVELS[5], ch[4], ch[5], 2, level_root=True)
    # DLAUp (lore_dla_34.py:113-137) over channels[2:] = [64,128,256,512], scales [1,2,4,8]
    channels = list(ch[2:])
    in_channels = list(channels)
    scales = np.array([1, 2, 4, 8], dtype=int)
    for i in range(len(channels) - 1):
        j = -i - 2
        _ida(rng, sd, f"dla_up.ida_{i}", channels[j], in_channels[j:], scales[j:] // scales[j], perturb_up)
        scales[j + 1:] = scales[j]
        in_channels[j + 1:] = [channels[j] for _ in channels[j + 1:]]
    _ida(rng, sd, "ida_up", ch[2], list(ch[2:5]), [1, 2, 4], perturb_up)
    for head, classes in LORE_HEADS:
        sd[f"{head}.0.weight"] = _conv(rng, 256, ch[2], 3, 3)
        sd[f"{head}.0.bias"] = _b(rng, 256)
        sd[f"{head}.2.weight"] = _conv(rng, classes, 256, 1, 1, gain=1.0)
        sd[f"{head}.2.bias"] = _b(rng, classes) if head != "hm" else np.full(classes, -2.19, np.float32)
    return sd


def lore_processor_state_dict(seed: int = 0, layers: int = 4, stacking_layers: int = 4) -> "OrderedDict[str, np.ndarray]":
    """Keys / shapes of `LoreProcessModel` (reference model/lore/lore_processor.py:399-514), wtw configuration
    (tsfm_layers = stacking_layers = 4, d = 256, 8 heads, FFN 2048)."""
    rng = np.random.Generator(np.random.PCG64(3000 + seed))
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    d, dff = 256, 2048

    def lin(p, cout, cin, gain=1.0):
        sd[p + ".weight"] = _lin(rng, cout, cin, gain)
        sd[p + ".bias"] = _b(rng, cout)

    def norm(p):
        sd[p + ".alpha"] = rng.uniform(0.5, 1.5, d).astype(np.float32)
        sd[p + ".bias"] = _b(rng, d)

    def transformer(p, cin, n_layers):
        lin(p + ".linear", d, cin)
        for L in range(n_layers):
            lp = f"{p}.encoder.layers.{L}"
            norm(lp + ".norm_1")
            norm(lp + ".norm_2")
            for n in ("q_linear", "v_linear", "k_linear"):
                lin(f"{lp}.attn.{n}", d, d)
            lin(lp + ".attn.out", d, d, 0.5)
            lin(lp + ".ff.linear_1", dff, d, 2.0)
            lin(lp + ".ff.linear_2", d, dff, 0.5)
        norm(p + ".encoder.norm")
        lin(p + ".decoder.linear.0", d, d, 2.0)
        lin(p + ".decoder.linear.2", 4, d, 2.0)

    lin("stacker.logi_encoder.0", d, 4, 2.0)
    lin("stacker.logi_encoder.2", d, d, 2.0)
    transformer("stacker.tsfm", 2 * d, stacking_layers)
    transformer("tsfm_axis", d, layers)
    sd["x_position_embeddings.weight"] = _b(rng, 256 * d, 1.0).reshape(256, d)
    sd["y_position_embeddings.weight"] = _b(rng, 256 * d, 1.0).reshape(256, d)
    return sd


# --------------------------------------------------------------------------- planted Lore head maps
def lore_planted_maps(index: int, h: int = 128, w: int = 128, feat_dim: int = 256, with_feat: bool = True):
    """Head outputs of a Lore detector for a planted ruled table (numpy only, deterministic): `hm` [2,h,w] AFTER the
    sigmoid (class 0 = cell centres, class 1 = corner points), `reg` [2,h,w], `wh` [8,h,w], `st` [8,h,w] and, when
    with_feat, random `ax` / `cr` [feat_dim,h,w].  A jittered rows x cols grid: every cell gets a centre peak whose
    `wh` points (noisily) at its four corners; every grid intersection gets a corner peak whose `st` spans a small
    cross reaching into the adjacent cells.  Some peaks are weak (below the 0.2 / 0.3 gates, or weak enough that the
    x0.4 penalty drops them), some corners are missing, and a low random floor creates thousands of sub-threshold
    local maxima, so every branch of the reference's decode is exercised."""
    rng = np.random.default_rng(777000 + index)
    hm = (rng.random((2, h, w)) * 0.12).astype(np.float32)
    reg = rng.random((2, h, w)).astype(np.float32)
    wh = (rng.standard_normal((8, h, w)) * 2).astype(np.float32)
    st = (rng.standard_normal((8, h, w)) * 2).astype(np.float32)
    rows, cols = int(rng.integers(3, 9)), int(rng.integers(3, 8))
    x0, y0 = rng.uniform(4, 0.15 * w), rng.uniform(4, 0.15 * h)
    x1, y1 = rng.uniform(0.8 * w, w - 5), rng.uniform(0.8 * h, h - 5)
    gx = np.sort(np.concatenate([[x0, x1], rng.uniform(x0 + 6, x1 - 6, cols - 1)]))
    gy = np.sort(np.concatenate([[y0, y1], rng.uniform(y0 + 6, y1 - 6, rows - 1)]))
    skew = rng.uniform(-0.04, 0.04)
    px = lambda r, c: np.float32(gx[c] + skew * (gy[r] - y0) + rng.normal(0, 0.15))
    py = lambda r, c: np.float32(gy[r] + skew * (gx[c] - x0) + rng.normal(0, 0.15))
    P = np.array([[(px(r, c), py(r, c)) for c in range(cols + 1)] for r in range(rows + 1)], np.float32)

    def plant(cls, x, y, score):
        ix, iy = int(np.floor(x)), int(np.floor(y))
        if not (1 <= ix < w - 1 and 1 <= iy < h - 1):
            return None
        hm[cls, iy - 1:iy + 2, ix - 1:ix + 2] = np.minimum(hm[cls, iy - 1:iy + 2, ix - 1:ix + 2], np.float32(score * 0.5))
        hm[cls, iy, ix] = np.float32(score)
        reg_here = reg[:, iy, ix]
        return ix, iy, np.float32(ix) + reg_here[0], np.float32(iy) + reg_here[1]

    for r in range(rows):
        for c in range(cols):
            if rng.random() < 0.05:
                continue  # undetected cell
            quad = np.array([P[r, c], P[r, c + 1], P[r + 1, c + 1], P[r + 1, c]], np.float32)
            cx, cy = quad.mean(0)
            score = rng.uniform(0.22, 0.98) if rng.random() < 0.85 else rng.uniform(0.1, 0.2)
            got = plant(0, cx, cy, score)
            if got is None:
                continue
            ix, iy, fx, fy = got
            noisy = quad + rng.normal(0, 0.6, quad.shape).astype(np.float32)
            wh[0::2, iy, ix] = fx - noisy[:, 0]
            wh[1::2, iy, ix] = fy - noisy[:, 1]
    for r in range(rows + 1):
        for c in range(cols + 1):
            if rng.random() < 0.12:
                continue  # missing corner -> some cells end with count <= 2
            score = rng.uniform(0.32, 0.97) if rng.random() < 0.9 else rng.uniform(0.15, 0.3)
            got = plant(1, P[r, c, 0], P[r, c, 1], score)
            if got is None:
                continue
            ix, iy, fx, fy = got
            arm = rng.uniform(1.5, 4.0)
            cross = np.array([[-arm, -arm], [arm, -arm], [arm, arm], [-arm, arm]], np.float32) + rng.normal(0, 0.3, (4, 2)).astype(np.float32)
            st[0::2, iy, ix] = -cross[:, 0]
            st[1::2, iy, ix] = -cross[:, 1]
    out = {"hm": hm, "reg": reg, "wh": wh, "st": st}
    if with_feat:
        out["ax"] = rng.standard_normal((feat_dim, h, w)).astype(np.float32)
        out["cr"] = rng.standard_normal((feat_dim, h, w)).astype(np.float32)
    return out


# --------------------------------------------------------------------------- planted PicoDet head outputs
PICODET_STRIDES = (8, 16, 32, 64)


def picodet_planted_outputs(index: int, num_classes: int = 5, in_h: int = 800, in_w: int = 608, reg_max: int = 7,
                            n_objects: int = 12):
    """Outputs of a PicoDet head (`export_post_process=False` contract, reference picodet/pico_head.py:1130-1138) for a
    page with planted layout regions: per level l, `scores[l]` fp32 [1, HW_l, C] (sigmoid class scores) and
    `boxes[l]` fp32 [1, HW_l, 4*(reg_max+1)] (raw DFL logits).  Anchors near the centre of a planted region get a high
    score for its class and DFL logits peaked (softly, with noise) at distance/stride, so several anchors per object
    predict overlapping boxes and the per-class NMS has real work; the rest is low-score noise."""
    rng = np.random.default_rng(31000 + index)
    scores, boxes = [], []
    objs = []
    for _ in range(n_objects):
        bw, bh = rng.uniform(60, in_w * 0.8), rng.uniform(30, in_h * 0.4)
        cx, cy = rng.uniform(bw / 2, in_w - bw / 2), rng.uniform(bh / 2, in_h - bh / 2)
        objs.append((cx - bw / 2, cy - bh / 2, cx + bw / 2, cy + bh / 2, int(rng.integers(0, num_classes)), rng.uniform(0.55, 0.97)))
    for stride in PICODET_STRIDES:
        fh, fw = int(np.ceil(in_h / stride)), int(np.ceil(in_w / stride))
        sc = (rng.random((fh * fw, num_classes)) * 0.3).astype(np.float32) ** 2
        bx = (rng.standard_normal((fh * fw, 4 * (reg_max + 1))) * 0.5).astype(np.float32)
        ys, xs = np.divmod(np.arange(fh * fw), fw)
        pcx, pcy = (xs + 0.5) * stride, (ys + 0.5) * stride
        for x0, y0, x1, y1, cls, peak in objs:
            d = np.stack([pcx - x0, pcy - y0, x1 - pcx, y1 - pcy], 1) / stride  # l, t, r, b in stride units
            ok = (d.min(1) > 0.3) & (d.max(1) < reg_max - 0.2)
            # only anchors near the object's centre fire (like a trained head's centre prior)
            near = (np.abs(pcx - (x0 + x1) / 2) < 0.2 * (x1 - x0)) & (np.abs(pcy - (y0 + y1) / 2) < 0.2 * (y1 - y0))
            sel = np.flatnonzero(ok & near)
            for a in sel:
                sc[a, cls] = np.float32(np.clip(peak - rng.uniform(0, 0.25), 0.05, 0.99))
                for side in range(4):
                    t = d[a, side] + rng.normal(0, 0.08)
                    lo = int(np.floor(t))
                    logits = np.full(reg_max + 1, -4.0)
                    logits[np.clip(lo, 0, reg_max)] = 4.0 + np.log(max(1e-3, 1 - (t - lo)))
                    logits[np.clip(lo + 1, 0, reg_max)] = 4.0 + np.log(max(1e-3, t - lo))
                    bx[a, side * (reg_max + 1):(side + 1) * (reg_max + 1)] = (logits + rng.normal(0, 0.1, reg_max + 1)).astype(np.float32)
        scores.append(sc[None])
        boxes.append(bx[None])
    return scores, boxes


# --------------------------------------------------------------------------- PicoDet (LCNet-x1.0 + CSP-PAN + PicoHead)
LCNET_CONFIG = {  # reference picodet/lcnet.py:25-46: k, in_c, out_c, stride, use_se
    "blocks2": [[3, 16, 32, 1, False]],
    "blocks3": [[3, 32, 64, 2, False], [3, 64, 64, 1, False]],
    "blocks4": [[3, 64, 128, 2, False], [3, 128, 128, 1, False]],
    "blocks5": [[3, 128, 256, 2, False]] + [[5, 256, 256, 1, False]] * 5,
    "blocks6": [[5, 256, 512, 2, True], [5, 512, 512, 1, True]],
}
PICO_NECK_CH, PICO_NECK_K, PICO_HEAD_CONVS, PICO_REG_MAX = 128, 5, 4, 7


def picodet_state_dicts(seed: int = 0, num_classes: int = 5):
    """Seeded weights for the three PicoDet modules with the reference's key names / shapes:
    LCNet(scale=1.0, feature_maps=[3,4,5]) (picodet/lcnet.py:159), CSPPAN(in=[128,256,512], out=128, kernel 5,
    num_features=4, depthwise, hard_swish) (picodet/csp_pan.py:233), PicoHead(PicoFeat(128 -> 128, 4 strides, 4 convs,
    share_cls_reg, use_se), fpn_stride [8,16,32,64], reg_max 7) (picodet/pico_head.py:37-167, 972-1160) -- the
    hyper-parameters of PaddleDetection's picodet_lcnet_x1_0_layout.  Returns (backbone, neck, head)."""
    rng = np.random.Generator(np.random.PCG64(4000 + seed))

    def conv_bn(sd, p, cout, cin, k, groups=1, conv="conv", bn="bn", gain=1.0):
        sd[f"{p}.{conv}.weight"] = _conv(rng, cout, cin // groups, k, k, gain)
        _bn(rng, sd, f"{p}.{bn}", cout)

    bb: "OrderedDict[str, np.ndarray]" = OrderedDict()
    conv_bn(bb, "conv1", 16, 3, 3)
    for name, cfg in LCNET_CONFIG.items():
        for i, (k, cin, cout, s, se) in enumerate(cfg):
            p = f"{name}.{i}"
            conv_bn(bb, p + ".dw_conv", cin, cin, k, groups=cin)
            if se:
                bb[p + ".se.conv1.weight"] = _conv(rng, cin // 4, cin, 1, 1, 1.0)
                bb[p + ".se.conv1.bias"] = _b(rng, cin // 4)
                bb[p + ".se.conv2.weight"] = _conv(rng, cin, cin // 4, 1, 1, 1.0)
                bb[p + ".se.conv2.bias"] = _b(rng, cin, 0.5)
            conv_bn(bb, p + ".pw_conv", cout, cin, 1)
    nk: "OrderedDict[str, np.ndarray]" = OrderedDict()
    c = PICO_NECK_CH

    def dp(p, ch, k):  # DPModule
        nk[p + ".dwconv.weight"] = _conv(rng, ch, 1, k, k, 1.0)
        _bn(rng, nk, p + ".bn1", ch)
        nk[p + ".pwconv.weight"] = _conv(rng, ch, ch, 1, 1, 1.0)
        _bn(rng, nk, p + ".bn2", ch)

    def csp(p):
        mid = c // 2
        conv_bn(nk, p + ".main_conv", mid, 2 * c, 1)
        conv_bn(nk, p + ".short_conv", mid, 2 * c, 1)
        conv_bn(nk, p + ".final_conv", c, 2 * mid, 1)
        conv_bn(nk, p + ".blocks.0.conv1", mid, mid, 1)
        dp(p + ".blocks.0.conv2", mid, PICO_NECK_K)

    for i, cin in enumerate((128, 256, 512)):
        conv_bn(nk, f"conv_t.convs.{i}", c, cin, 1)
    dp("first_top_conv", c, PICO_NECK_K)
    dp("second_top_conv", c, PICO_NECK_K)
    csp("top_down_blocks.0")
    csp("top_down_blocks.1")
    dp("downsamples.0", c, PICO_NECK_K)
    dp("downsamples.1", c, PICO_NECK_K)
    csp("bottom_up_blocks.0")
    csp("bottom_up_blocks.1")
    hd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    for lvl in range(4):  # PicoSE: built, but with share_cls_reg its output is never read (pico_head.py:1117-1124)
        hd[f"conv_feat.se.{lvl}.fc.weight"] = _conv(rng, c, c, 1, 1, 1.0)
        hd[f"conv_feat.se.{lvl}.fc.bias"] = _b(rng, c)
        conv_bn(hd, f"conv_feat.se.{lvl}.conv", c, c, 1, bn="norm")
    for lvl in range(4):
        for i in range(PICO_HEAD_CONVS):
            conv_bn(hd, f"conv_feat.cls_conv_dw{lvl}_{i}", c, c, 5, groups=c, bn="norm")
            conv_bn(hd, f"conv_feat.cls_conv_pw{lvl}_{i}", c, c, 1, bn="norm")
    for lvl in range(4):
        hd[f"head_cls{lvl}.weight"] = _conv(rng, num_classes + 4 * (PICO_REG_MAX + 1), c, 1, 1, 1.0)
        bias = _b(rng, num_classes + 4 * (PICO_REG_MAX + 1), 0.5)
        bias[:num_classes] -= 2.0  # class prior: most anchors are background
        hd[f"head_cls{lvl}.bias"] = bias
    return bb, nk, hd


# --------------------------------------------------------------------------- CenterNet table structure (DLA-34, plain IDA-up)
CENTERNET_HEADS = (("hm", 2), ("v2c", 8), ("c2v", 8), ("reg", 2))


def centernet_dla34_state_dict(seed: int = 0, perturb_up: bool = True) -> "OrderedDict[str, np.ndarray]":
    """Keys / shapes of the reference CenterNet `DLASeg()` (center_net/modeling_centernet.py:601-668: dla34 base, DLAUp of
    plain IDAUp blocks :508-599 with node_kernel 3, heads hm 2 / v2c 8 / c2v 8 / reg 2 with head_conv 256)."""
    rng = np.random.Generator(np.random.PCG64(5000 + seed))
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    ch = DLA_CHANNELS
    sd["base.base_layer.0.weight"] = _conv(rng, ch[0], 3, 7, 7)
    _bn(rng, sd, "base.base_layer.1", ch[0])
    sd["base.level0.0.weight"] = _conv(rng, ch[0], ch[0], 3, 3)
    _bn(rng, sd, "base.level0.1", ch[0])
    sd["base.level1.0.weight"] = _conv(rng, ch[1], ch[0], 3, 3)
    _bn(rng, sd, "base.level1.1", ch[1])
    _dla_tree(rng, sd, "base.level2", DLA_LEVELS[2], ch[1], ch[2], 2, level_root=False)
    _dla_tree(rng, sd, "base.level3", DLA_LEVELS[3], ch[2], ch[3], 2, level_root=True)
    _dla_tree(rng, sd, "base.level4", DLA_LEVELS[4], ch[3], ch[4], 2, level_root=True)
    _dla_tree(rng, sd, "base.level5", DLA_LEVELS[5], ch[4], ch[5], 2, level_root=True)
    channels = list(ch[2:])
    in_channels = list(channels)
    scales = np.array([1, 2, 4, 8], dtype=int)
    for i in range(len(channels) - 1):
        j = -i - 2
        o, cin, up_f = channels[j], in_channels[j:], scales[j:] // scales[j]
        p = f"dla_up.ida_{i}"
        for k, c in enumerate(cin):
            if c != o:
                sd[f"{p}.proj_{k}.0.weight"] = _conv(rng, o, c, 1, 1, gain=0.5)
                _bn(rng, sd, f"{p}.proj_{k}.1", o)
            if int(up_f[k]) != 1:
                w = _bilinear_up(o, int(up_f[k]))
                if perturb_up:
                    w = (w * rng.uniform(0.8, 1.2, w.shape)).astype(np.float32)
                sd[f"{p}.up_{k}.weight"] = w
        for k in range(1, len(cin)):
            sd[f"{p}.node_{k}.0.weight"] = _conv(rng, o, 2 * o, 3, 3, gain=0.5)  # keeps the un-normalised synthetic activations O(1)
            _bn(rng, sd, f"{p}.node_{k}.1", o)
        scales[j + 1:] = scales[j]
        in_channels[j + 1:] = [channels[j] for _ in channels[j + 1:]]
    for head, classes in CENTERNET_HEADS:
        sd[f"{head}.0.weight"] = _conv(rng, 256, ch[2], 3, 3, gain=0.5)
        sd[f"{head}.0.bias"] = _b(rng, 256)
        sd[f"{head}.2.weight"] = _conv(rng, classes, 256, 1, 1, gain=1.0)
        sd[f"{head}.2.bias"] = _b(rng, classes) if head != "hm" else np.full(classes, -2.19, np.float32)
    return sd


# --------------------------------------------------------------------------- PP-OCRv4 rec (PPLCNetV3-0.95 + SVTR neck + CTC)
PP_REC_CONFIG = {  # PaddleOCR rec_lcnetv3.py NET_CONFIG_rec: k, in_c, out_c, stride, use_se
    "blocks2": [[3, 16, 32, 1, False]],
    "blocks3": [[3, 32, 64, 1, False], [3, 64, 64, 1, False]],
    "blocks4": [[3, 64, 128, (2, 1), False], [3, 128, 128, 1, False]],
    "blocks5": [[3, 128, 256, (1, 2), False], [5, 256, 256, 1, False], [5, 256, 256, 1, False], [5, 256, 256, 1, False],
                [5, 256, 256, 1, False]],
    "blocks6": [[5, 256, 512, (2, 1), True], [5, 512, 512, 1, True], [5, 512, 512, (2, 1), False], [5, 512, 512, 1, False]],
}


def pp_rec_ch(c: int, scale: float = 0.95) -> int:
    """make_divisible(c * scale, 16) of rec_lcnetv3.py."""
    v = c * scale
    new_v = max(16, int(v + 8) // 16 * 16)
    return new_v + 16 if new_v < 0.9 * v else new_v


def pp_ocrv4_rec_state_dict(seed: int = 0, n_class: int = 97) -> "OrderedDict[str, np.ndarray]":
    """Seeded weights of the deploy-form PP-OCRv4 recogniser (keys: oracle/pp_rec_ref.py).  The LearnableAffineBlock scalars are
    drawn away from their (1, 0) initial values so that every affine is exercised."""
    rng = np.random.Generator(np.random.PCG64(seed + 4040))
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()

    def lab(p):
        sd[p + ".scale"] = rng.uniform(0.8, 1.25, 1).astype(np.float32)
        sd[p + ".bias"] = (rng.standard_normal(1) * 0.05).astype(np.float32)

    def rep(p, cin, cout, k, groups):
        sd[p + ".reparam_conv.weight"] = _conv(rng, cout, cin // groups, k, k)
        sd[p + ".reparam_conv.bias"] = _b(rng, cout)
        lab(p + ".lab")
        lab(p + ".act.lab")

    sd["backbone.conv1.conv.weight"] = _conv(rng, 16, 3, 3, 3)
    _bn(rng, sd, "backbone.conv1.bn", 16)
    for name, cfg in PP_REC_CONFIG.items():
        for i, (k, cin, cout, s, se) in enumerate(cfg):
            ci, co = pp_rec_ch(cin), pp_rec_ch(cout)
            p = f"backbone.{name}.{i}"
            rep(p + ".dw_conv", ci, ci, k, ci)
            if se:
                sd[p + ".se.conv1.weight"] = _conv(rng, ci // 4, ci, 1, 1)
                sd[p + ".se.conv1.bias"] = _b(rng, ci // 4)
                sd[p + ".se.conv2.weight"] = _conv(rng, ci, ci // 4, 1, 1)
                sd[p + ".se.conv2.bias"] = _b(rng, ci)
            rep(p + ".pw_conv", ci, co, 1, 1)
    cb, d = pp_rec_ch(512), 120
    p = "head.ctc_encoder.encoder"
    for name, cin, cout, kw in (("conv1", cb, cb // 8, 3), ("conv2", cb // 8, d, 1), ("conv3", d, cb, 1), ("conv4", 2 * cb, cb // 8, 3),
                                ("conv1x1", cb // 8, d, 1)):
        sd[f"{p}.{name}.conv.weight"] = _conv(rng, cout, cin, 1, kw)
        _bn(rng, sd, f"{p}.{name}.norm", cout)
    for i in range(2):
        q = f"{p}.svtr_block.{i}"
        _ln(rng, sd, q + ".norm1", d)
        sd[q + ".mixer.qkv.weight"], sd[q + ".mixer.qkv.bias"] = _lin(rng, 3 * d, d), _b(rng, 3 * d)
        sd[q + ".mixer.proj.weight"], sd[q + ".mixer.proj.bias"] = _lin(rng, d, d), _b(rng, d)
        _ln(rng, sd, q + ".norm2", d)
        sd[q + ".mlp.fc1.weight"], sd[q + ".mlp.fc1.bias"] = _lin(rng, 2 * d, d), _b(rng, 2 * d)
        sd[q + ".mlp.fc2.weight"], sd[q + ".mlp.fc2.bias"] = _lin(rng, d, 2 * d), _b(rng, d)
    _ln(rng, sd, p + ".norm", d)
    sd["head.ctc_head.fc.weight"], sd["head.ctc_head.fc.bias"] = _lin(rng, n_class, d, gain=4.0), _b(rng, n_class)
    return sd


# --------------------------------------------------------------------------- PP-OCRv4 mobile detector (PPLCNetV3-0.75 + RSE-FPN + DBHead)
PP_DET_CONFIG = {  # PaddleOCR rec_lcnetv3.py NET_CONFIG_det: k, in_c, out_c, stride, use_se
    "blocks2": [[3, 16, 32, 1, False]],
    "blocks3": [[3, 32, 64, 2, False], [3, 64, 64, 1, False]],
    "blocks4": [[3, 64, 128, 2, False], [3, 128, 128, 1, False]],
    "blocks5": [[3, 128, 256, 2, False], [5, 256, 256, 1, False], [5, 256, 256, 1, False], [5, 256, 256, 1, False],
                [5, 256, 256, 1, False]],
    "blocks6": [[5, 256, 512, 2, True], [5, 512, 512, 1, True], [5, 512, 512, 1, False], [5, 512, 512, 1, False]],
}
PP_DET_SCALE, PP_DET_MV_C, PP_DET_FPN = 0.75, (16, 24, 56, 480), 96


def pp_ocrv4_det_state_dict(seed: int = 0) -> "OrderedDict[str, np.ndarray]":
    """Seeded weights of the deploy-form PP-OCRv4 mobile detector (keys: oracle/pp_det_ref.py)."""
    rng = np.random.Generator(np.random.PCG64(seed + 4141))
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    ch = lambda c: pp_rec_ch(c, PP_DET_SCALE)  # noqa: E731

    def lab(p):
        sd[p + ".scale"] = rng.uniform(0.8, 1.25, 1).astype(np.float32)
        sd[p + ".bias"] = (rng.standard_normal(1) * 0.05).astype(np.float32)

    def rep(p, cin, cout, k, groups, stride):
        sd[p + ".reparam_conv.weight"] = _conv(rng, cout, cin // groups, k, k)
        sd[p + ".reparam_conv.bias"] = _b(rng, cout)
        lab(p + ".lab")
        if stride != 2:
            lab(p + ".act.lab")

    def se(p, c):
        sd[p + ".conv1.weight"] = _conv(rng, c // 4, c, 1, 1)
        sd[p + ".conv1.bias"] = _b(rng, c // 4)
        sd[p + ".conv2.weight"] = _conv(rng, c, c // 4, 1, 1)
        sd[p + ".conv2.bias"] = _b(rng, c)

    sd["backbone.conv1.conv.weight"] = _conv(rng, 16, 3, 3, 3)
    _bn(rng, sd, "backbone.conv1.bn", 16)
    tap_in = []
    for name, cfg in PP_DET_CONFIG.items():
        for i, (k, cin, cout, s, use_se) in enumerate(cfg):
            ci, co = ch(cin), ch(cout)
            p = f"backbone.{name}.{i}"
            rep(p + ".dw_conv", ci, ci, k, ci, s)
            if use_se:
                se(p + ".se", ci)
            rep(p + ".pw_conv", ci, co, 1, 1, 1)
        if name != "blocks2":
            tap_in.append(co)
    taps = [int(c * PP_DET_SCALE) for c in PP_DET_MV_C]
    for i, (ci, co) in enumerate(zip(tap_in, taps)):
        sd[f"backbone.layer_list.{i}.weight"] = _conv(rng, co, ci, 1, 1)
        sd[f"backbone.layer_list.{i}.bias"] = _b(rng, co)
    f = PP_DET_FPN
    for i, c in enumerate(taps):
        # linear stages (no activation follows): unit gain keeps the synthetic logits in a few units instead of saturating the sigmoid
        sd[f"neck.ins_conv.{i}.in_conv.weight"] = _conv(rng, f, c, 1, 1, gain=0.5)
        se(f"neck.ins_conv.{i}.se_block", f)
        sd[f"neck.inp_conv.{i}.in_conv.weight"] = _conv(rng, f // 4, f, 3, 3, gain=0.25)
        se(f"neck.inp_conv.{i}.se_block", f // 4)
    p = "head.binarize"
    sd[p + ".conv1.weight"] = _conv(rng, f // 4, f, 3, 3, gain=1.0)
    _bn(rng, sd, p + ".conv_bn1", f // 4)
    sd[p + ".conv2.weight"] = (rng.standard_normal((f // 4, f // 4, 2, 2)) * np.sqrt(1.0 / (f // 4))).astype(np.float32)
    sd[p + ".conv2.bias"] = _b(rng, f // 4)
    _bn(rng, sd, p + ".conv_bn2", f // 4)
    sd[p + ".conv3.weight"] = (rng.standard_normal((f // 4, 1, 2, 2)) * np.sqrt(1.0 / (f // 4))).astype(np.float32)
    sd[p + ".conv3.bias"] = _b(rng, 1)
    return sd


# --------------------------------------------------------------------------- CRNN (crnn/modeling_crnn.py)
CRNN_LABELS = 7644


def crnn_state_dict(seed: int = 0, num_labels: int = CRNN_LABELS) -> "OrderedDict[str, np.ndarray]":
    """Seeded weights with the keys / shapes of the reference CRNN module (crnn/modeling_crnn.py:36-88)."""
    rng = np.random.Generator(np.random.PCG64(seed + 6060))
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()

    def conv(p, cout, cin, kh, kw):
        sd[p + ".weight"] = _conv(rng, cout, cin, kh, kw)
        sd[p + ".bias"] = _b(rng, cout)

    conv("conv0.0", 64, 1, 3, 3)
    _bn(rng, sd, "conv0.1", 64)
    conv("conv1.0", 128, 64, 3, 3)
    _bn(rng, sd, "conv1.1", 128)
    conv("conv2.0", 256, 128, 3, 3)
    _bn(rng, sd, "conv2.1", 256)
    conv("conv2.3", 256, 256, 3, 3)
    _bn(rng, sd, "conv2.4", 256)
    conv("conv3.0", 512, 256, 3, 3)
    _bn(rng, sd, "conv3.1", 512)
    conv("conv3.3", 512, 512, 3, 3)
    _bn(rng, sd, "conv3.4", 512)
    conv("conv4.0", 512, 512, 2, 1)
    _bn(rng, sd, "conv4.1", 512)
    for layer, (nin, nout) in enumerate(((512, 256), (256, 512))):
        k = 1.0 / np.sqrt(256.0)  # nn.LSTM's own init range
        for suffix in ("", "_reverse"):
            sd[f"rnn.{layer}.rnn.weight_ih_l0{suffix}"] = rng.uniform(-k, k, (1024, nin)).astype(np.float32) * np.float32(np.sqrt(256.0 / nin) * 1.5)
            sd[f"rnn.{layer}.rnn.weight_hh_l0{suffix}"] = rng.uniform(-k, k, (1024, 256)).astype(np.float32) * np.float32(1.5)
            sd[f"rnn.{layer}.rnn.bias_ih_l0{suffix}"] = rng.uniform(-k, k, 1024).astype(np.float32)
            sd[f"rnn.{layer}.rnn.bias_hh_l0{suffix}"] = rng.uniform(-k, k, 1024).astype(np.float32)
        sd[f"rnn.{layer}.embedding.weight"] = _lin(rng, nout, 512, gain=2.0)
        sd[f"rnn.{layer}.embedding.bias"] = _b(rng, nout)
    sd["cls.weight"] = _lin(rng, num_labels, 512, gain=8.0)
    return sd


# --------------------------------------------------------------------------- PULC classifiers (PP-LCNet x1.0, cls/cls_pp_lcnet.py)
def pplcnet_cls_state_dict(seed: int = 0, class_num: int = 4) -> "OrderedDict[str, np.ndarray]":
    """Seeded weights with the keys of the reference PPLCNet module (cls/cls_pp_lcnet.py:164-293, scale 1.0, class_expand 1280)."""
    from .pplcnet_graph import NET_CONFIG

    rng = np.random.Generator(np.random.PCG64(seed + 5050))
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    sd["conv1.conv.weight"] = _conv(rng, 16, 3, 3, 3)
    _bn(rng, sd, "conv1.bn", 16)
    for name, cfg in NET_CONFIG.items():
        for i, (k, ci, co, s, se) in enumerate(cfg):
            p = f"{name}.{i}"
            sd[p + ".dw_conv.conv.weight"] = _conv(rng, ci, 1, k, k)
            _bn(rng, sd, p + ".dw_conv.bn", ci)
            if se:
                sd[p + ".se.conv1.weight"], sd[p + ".se.conv1.bias"] = _conv(rng, ci // 4, ci, 1, 1), _b(rng, ci // 4)
                sd[p + ".se.conv2.weight"], sd[p + ".se.conv2.bias"] = _conv(rng, ci, ci // 4, 1, 1), _b(rng, ci)
            sd[p + ".pw_conv.conv.weight"] = _conv(rng, co, ci, 1, 1)
            _bn(rng, sd, p + ".pw_conv.bn", co)
    sd["last_conv.weight"] = _conv(rng, 1280, 512, 1, 1)
    sd["fc.weight"], sd["fc.bias"] = _lin(rng, class_num, 1280, gain=400.0), _b(rng, class_num)
    return sd


# --------------------------------------------------------------------------- Lore wireless (ResNet-18 key-point detector)
LORE_R18_PLANES = (64, 128, 256, 256)  # layer1..4, every one entered with stride 2 (lore/lore_detector.py:180-187)
LORE_R18_HEAD_CONV = 64


def lore_resnet18_state_dict(seed: int = 0) -> "OrderedDict[str, np.ndarray]":
    """Keys / shapes of `LoreDetectModel` (reference model/lore/lore_detector.py:148-389): 7x7 stem, four 2-block stages
    (BasicBlock convs WITH bias, :75-79), four ConvTranspose 4x4 s2 up-steps with 1x1 `adaption` laterals, six heads of
    head_conv = 64 (`reg`: conv3x3 + 1x1; the others: four conv3x3 + 1x1)."""
    rng = np.random.Generator(np.random.PCG64(7000 + seed))
    sd: "OrderedDict[str, np.ndarray]" = OrderedDict()
    sd["conv1.weight"] = _conv(rng, 64, 3, 7, 7)
    _bn(rng, sd, "bn1", 64)
    inpl = 64
    for L, planes in enumerate(LORE_R18_PLANES, start=1):
        for B in range(2):
            p = f"layer{L}.{B}"
            sd[p + ".conv1.weight"] = _conv(rng, planes, inpl, 3, 3)
            sd[p + ".conv1.bias"] = _b(rng, planes)
            _bn(rng, sd, p + ".bn1", planes)
            sd[p + ".conv2.weight"] = _conv(rng, planes, planes, 3, 3, gain=1.0)
            sd[p + ".conv2.bias"] = _b(rng, planes)
            _bn(rng, sd, p + ".bn2", planes)
            if B == 0:
                sd[p + ".downsample.0.weight"] = _conv(rng, planes, inpl, 1, 1, gain=1.0)
                _bn(rng, sd, p + ".downsample.1", planes)
            inpl = planes
    for name, cin in (("adaption3", 256), ("adaption2", 128), ("adaption1", 64), ("adaption0", 64), ("adaptionU1", 256)):
        sd[name + ".weight"] = _conv(rng, 256, cin, 1, 1, gain=1.0)
    for i in range(1, 5):  # ConvTranspose2d weight [Cin, Cout, 4, 4]; every output pixel sums 2 x 2 taps x 256 channels
        sd[f"deconv_layers{i}.0.weight"] = (rng.standard_normal((256, 256, 4, 4)) * np.sqrt(2.0 / (256 * 4))).astype(np.float32)
        _bn(rng, sd, f"deconv_layers{i}.1", 256)
    hc = LORE_R18_HEAD_CONV
    for head, classes in sorted(LORE_HEADS):
        sd[f"{head}.0.weight"] = _conv(rng, hc, 256, 3, 3)
        sd[f"{head}.0.bias"] = _b(rng, hc)
        last = 2
        if head != "reg":
            for j in (2, 4, 6):
                sd[f"{head}.{j}.weight"] = _conv(rng, hc, hc, 3, 3, gain=1.4)
                sd[f"{head}.{j}.bias"] = _b(rng, hc)
            last = 8
        sd[f"{head}.{last}.weight"] = _conv(rng, classes, hc, 1, 1, gain=0.25)
        sd[f"{head}.{last}.bias"] = _b(rng, classes) if head != "hm" else np.full(classes, -2.19, np.float32)
    return sd
