"""Plain-PyTorch fp32 restatement of the reference Lore detector `get_dla_dcn(34, heads)` (TEST ORACLE, see
oracle/__init__.py; never imported by the product).

Follows, in functional form over a numpy/torch state_dict:
  * DLA-34 base          center_net/modeling_centernet.py: BasicBlock :34-71, Root :159-183, Tree :186-287
                         (Tree ignores the `residual` it is handed and recomputes it from `bottom` :266-270),
                         DLA.forward :363-376 with levels [1,1,1,2,2,1], channels [16,32,64,128,256,512] (:383-387)
  * DCN                  lore/dcnv2.py:71-86: conv_offset_mask -> chunk(3) -> offset = cat(o1,o2) (= the first 18
                         channels, torchvision layout [dy_k, dx_k] per tap), mask = sigmoid(last 9) -> deform_conv2d
  * DeformConv/IDAUp/DLAUp/DLASeg   lore/lore_dla_34.py:65-85, 88-110 (up(proj(x)) then node(x + previous)),
                         113-137 (in-place mutation of `layers`), 140-190 (first_level 2, last_level 5, y[-1] -> heads)
Pinned against the reference module itself by tests/golden/lore_dla34_seed0.npz (oracle/gen_golden_lore.py).
`deform_conv2d_ref` restates torchvision's bilinear sampling rule (deform_conv2d_kernel.cpp bilinear_interpolate)
in plain torch and is checked against torchvision.ops.deform_conv2d in tests/test_oracle_cpu.py.
"""
from __future__ import annotations

from typing import Dict, List, Mapping

import numpy as np
import torch
import torch.nn.functional as F

from pdf_table_b200.synth import DLA_CHANNELS, DLA_LEVELS, LORE_HEADS


def _t(sd, k):
    v = sd[k]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))


def _bn(x, sd, p, eps=1e-5):
    return F.batch_norm(x, _t(sd, p + ".running_mean"), _t(sd, p + ".running_var"), _t(sd, p + ".weight"),
                        _t(sd, p + ".bias"), training=False, eps=eps)


def deform_conv2d_ref(x, offset, weight, bias, mask):
    """Modulated deformable 3x3 convolution, stride 1, pad 1, dilation 1, one offset group.
    x [N,C,H,W]; offset [N,18,H,W] = (dy_k, dx_k) per tap k = 3*i + j; mask [N,9,H,W]; weight [Co,C,3,3]."""
    n, c, h, w = x.shape
    ys = torch.arange(h, dtype=x.dtype).view(1, h, 1)
    xs = torch.arange(w, dtype=x.dtype).view(1, 1, w)
    cols = []
    flat = x.reshape(n, c, h * w)
    for k in range(9):
        i, j = divmod(k, 3)
        py = ys + (i - 1) + offset[:, 2 * k]
        px = xs + (j - 1) + offset[:, 2 * k + 1]
        inside = (py > -1) & (py < h) & (px > -1) & (px < w)
        y0 = torch.floor(py)
        x0 = torch.floor(px)
        ly, lx = py - y0, px - x0
        hy, hx = 1 - ly, 1 - lx
        y0, x0 = y0.long(), x0.long()
        y1, x1 = y0 + 1, x0 + 1

        def tap(yy, xx):
            ok = (yy >= 0) & (yy <= h - 1) & (xx >= 0) & (xx <= w - 1)
            idx = (yy.clamp(0, h - 1) * w + xx.clamp(0, w - 1)).view(n, 1, h * w).expand(n, c, h * w)
            return flat.gather(2, idx).view(n, c, h, w) * ok.view(n, 1, h, w)

        val = (hy * hx).unsqueeze(1) * tap(y0, x0) + (hy * lx).unsqueeze(1) * tap(y0, x1) + \
              (ly * hx).unsqueeze(1) * tap(y1, x0) + (ly * lx).unsqueeze(1) * tap(y1, x1)
        cols.append(val * inside.unsqueeze(1) * mask[:, k:k + 1])
    col = torch.stack(cols, 2)  # [N,C,9,H,W]
    out = torch.einsum("nckhw,ock->nohw", col, weight.reshape(weight.shape[0], c, 9))
    return out + bias.view(1, -1, 1, 1)


def _block(x, sd, p, stride, residual=None):
    if residual is None:
        residual = x
    out = F.relu(_bn(F.conv2d(x, _t(sd, p + ".conv1.weight"), stride=stride, padding=1), sd, p + ".bn1"))
    out = _bn(F.conv2d(out, _t(sd, p + ".conv2.weight"), padding=1), sd, p + ".bn2")
    return F.relu(out + residual)


def _tree(x, sd, p, levels, cin, cout, stride, level_root, children=None):
    children = [] if children is None else children
    bottom = F.max_pool2d(x, stride, stride) if stride > 1 else x
    residual = bottom
    if cin != cout:
        residual = _bn(F.conv2d(bottom, _t(sd, p + ".project.0.weight")), sd, p + ".project.1")
    if level_root:
        children.append(bottom)
    if levels == 1:
        x1 = _block(x, sd, p + ".tree1", stride, residual)
        x2 = _block(x1, sd, p + ".tree2", 1)
        cat = torch.cat([x2, x1] + children, 1)
        return F.relu(_bn(F.conv2d(cat, _t(sd, p + ".root.conv.weight")), sd, p + ".root.bn"))
    x1 = _tree(x, sd, p + ".tree1", levels - 1, cin, cout, stride, False)
    children.append(x1)
    return _tree(x1, sd, p + ".tree2", levels - 1, cout, cout, 1, False, children)


def dla34_base(sd, x) -> List[torch.Tensor]:
    ch = DLA_CHANNELS
    x = F.relu(_bn(F.conv2d(x, _t(sd, "base.base_layer.0.weight"), padding=3), sd, "base.base_layer.1"))
    y = []
    x = F.relu(_bn(F.conv2d(x, _t(sd, "base.level0.0.weight"), padding=1), sd, "base.level0.1"))
    y.append(x)
    x = F.relu(_bn(F.conv2d(x, _t(sd, "base.level1.0.weight"), stride=2, padding=1), sd, "base.level1.1"))
    y.append(x)
    for lvl in range(2, 6):
        x = _tree(x, sd, f"base.level{lvl}", DLA_LEVELS[lvl], ch[lvl - 1], ch[lvl], 2, lvl > 2)
        y.append(x)
    return y


def dcn_offsets(x, sd, p):
    om = F.conv2d(x, _t(sd, p + ".conv.conv_offset_mask.weight"), _t(sd, p + ".conv.conv_offset_mask.bias"), padding=1)
    return om[:, :18], torch.sigmoid(om[:, 18:27])


def deform_conv(x, sd, p, use_torchvision=True):
    offset, mask = dcn_offsets(x, sd, p)
    w, b = _t(sd, p + ".conv.weight"), _t(sd, p + ".conv.bias")
    if use_torchvision:
        from torchvision.ops import deform_conv2d
        out = deform_conv2d(x, offset=offset, weight=w, bias=b, stride=(1, 1), padding=(1, 1), dilation=(1, 1), mask=mask)
    else:
        out = deform_conv2d_ref(x, offset, w, b, mask)
    return F.relu(_bn(out, sd, p + ".actf.0"))


def _ida_up(layers, sd, p, startp, endp, up_f, tv):
    for i in range(startp + 1, endp):
        j = i - startp
        f = int(up_f[j])
        w = _t(sd, f"{p}.up_{j}.weight")
        t = deform_conv(layers[i], sd, f"{p}.proj_{j}", tv)
        t = F.conv_transpose2d(t, w, stride=f, padding=f // 2, groups=w.shape[0])
        layers[i] = deform_conv(t + layers[i - 1], sd, f"{p}.node_{j}", tv)


@torch.no_grad()
def lore_dla34_features(sd: Mapping[str, np.ndarray], x: torch.Tensor, use_torchvision: bool = True) -> torch.Tensor:
    """x fp32 [N,3,H,W] (H, W multiples of 32) -> the 64-channel stride-4 feature map y[-1] the heads read."""
    layers = dla34_base(sd, x.to(_t(sd, "base.base_layer.0.weight").dtype))[2:]  # first_level = 2: strides 4, 8, 16, 32
    scales = np.array([1, 2, 4, 8], dtype=int)
    out = [layers[-1]]
    for i in range(len(layers) - 1):
        j = -i - 2
        up_f = scales[j:] // scales[j]
        _ida_up(layers, sd, f"dla_up.ida_{i}", len(layers) - i - 2, len(layers), up_f, use_torchvision)
        out.insert(0, layers[-1])
        scales[j + 1:] = scales[j]
    y = [t.clone() for t in out[:3]]  # last_level - first_level = 3
    _ida_up(y, sd, "ida_up", 0, 3, [1, 2, 4], use_torchvision)
    return y[-1]


def lore_head(sd, feat, head):
    t = F.relu(F.conv2d(feat, _t(sd, f"{head}.0.weight"), _t(sd, f"{head}.0.bias"), padding=1))
    return F.conv2d(t, _t(sd, f"{head}.2.weight"), _t(sd, f"{head}.2.bias"))


@torch.no_grad()
def lore_dla34_forward(sd, x, use_torchvision: bool = True, heads=None) -> Dict[str, torch.Tensor]:
    """-> {'hm','st','wh','ax','cr','reg'} raw head outputs at stride 4 (hm NOT yet sigmoid-ed), plus 'feat'."""
    feat = lore_dla34_features(sd, x, use_torchvision)
    out = {"feat": feat}
    for head, _ in LORE_HEADS:
        if heads is None or head in heads:
            out[head] = lore_head(sd, feat, head)
    return out
