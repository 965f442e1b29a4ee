"""Plain-PyTorch fp32 restatement of the reference Lore `wireless` detector `LoreDetectModel` (TEST ORACLE, see
oracle/__init__.py; never imported by the product).

Follows, in functional form over a numpy / torch state_dict, lore/lore_detector.py:
  * BasicBlock :69-98 (both convs carry a bias; `downsample` = conv1x1 stride s + BN on the block's input)
  * LoreDetectModel.__init__ :153-285: conv 7x7 s2 + BN + ReLU, max-pool 3x3 s2 p1, layer1..4 = 2 blocks each with planes
    [64, 128, 256, 256], EVERY stage entered with stride 2 (:180-187 -- layer1 too, unlike torchvision's ResNet-18), so the
    stage outputs sit at strides 8 / 16 / 32 / 64 and x0 (the pooled stem) at stride 4
  * forward :353-389: x3_ = adaption3(x3) + up(x4); x2_ = adaption2(x2) + up(x3_); x1_ = adaption1(x1) + up(x2_);
    x0_ = adaptionU1(up(x1_) + adaption0(x0)); up = ConvTranspose2d(256, 256, 4, stride 2, padding 1, no bias) + BN + ReLU
    (:326-351); the six heads read x0_ (stride 4)
  * heads :240-285: `reg` = conv3x3(256 -> 64) + ReLU + conv1x1; every other head = four conv3x3 + ReLU (256 -> 64 -> 64 ->
    64 -> 64) + conv1x1
Pinned against the reference module itself by tests/golden/lore_resnet18_seed0.npz (oracle/gen_golden_lore_wireless.py).
"""
from __future__ import annotations

from typing import Dict, Mapping

import numpy as np
import torch
import torch.nn.functional as F

from pdf_table_b200.synth import LORE_HEADS


def _t(sd, k):
    v = sd[k]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))


def _bn(x, sd, p, eps=1e-5):
    return F.batch_norm(x, _t(sd, p + ".running_mean"), _t(sd, p + ".running_var"), _t(sd, p + ".weight"), _t(sd, p + ".bias"),
                        training=False, eps=eps)


def _block(x, sd, p, stride):
    out = F.relu(_bn(F.conv2d(x, _t(sd, p + ".conv1.weight"), _t(sd, p + ".conv1.bias"), stride=stride, padding=1), sd, p + ".bn1"))
    out = _bn(F.conv2d(out, _t(sd, p + ".conv2.weight"), _t(sd, p + ".conv2.bias"), padding=1), sd, p + ".bn2")
    res = x
    if (p + ".downsample.0.weight") in sd:
        res = _bn(F.conv2d(x, _t(sd, p + ".downsample.0.weight"), stride=stride), sd, p + ".downsample.1")
    return F.relu(out + res)


def _up(x, sd, i):
    return F.relu(_bn(F.conv_transpose2d(x, _t(sd, f"deconv_layers{i}.0.weight"), stride=2, padding=1), sd, f"deconv_layers{i}.1"))


def lore_resnet18_features(sd: Mapping[str, np.ndarray], x: torch.Tensor) -> Dict[str, torch.Tensor]:
    """x fp32 [N,3,H,W] (H, W multiples of 64) -> the intermediate maps by name; 'feat' = x0_ [N,256,H/4,W/4]."""
    t = {}
    t["c1"] = F.relu(_bn(F.conv2d(x, _t(sd, "conv1.weight"), stride=2, padding=3), sd, "bn1"))
    t["x0"] = F.max_pool2d(t["c1"], 3, 2, 1)
    y = t["x0"]
    for L in range(1, 5):
        y = _block(y, sd, f"layer{L}.0", 2)
        y = _block(y, sd, f"layer{L}.1", 1)
        t[f"x{L}"] = y
    lat = lambda name, v: F.conv2d(v, _t(sd, name + ".weight"))  # noqa: E731
    t["x3_"] = lat("adaption3", t["x3"]) + _up(t["x4"], sd, 1)
    t["x2_"] = lat("adaption2", t["x2"]) + _up(t["x3_"], sd, 2)
    t["x1_"] = lat("adaption1", t["x1"]) + _up(t["x2_"], sd, 3)
    t["x0s"] = _up(t["x1_"], sd, 4) + lat("adaption0", t["x0"])
    t["feat"] = lat("adaptionU1", t["x0s"])
    return t


def lore_r18_head(sd, feat, head, hidden: bool = False):
    y = feat
    last = 2 if head == "reg" else 8
    for j in range(0, last, 2):
        y = F.relu(F.conv2d(y, _t(sd, f"{head}.{j}.weight"), _t(sd, f"{head}.{j}.bias"), padding=1))
    if hidden:
        return y
    return F.conv2d(y, _t(sd, f"{head}.{last}.weight"), _t(sd, f"{head}.{last}.bias"))


@torch.no_grad()
def lore_resnet18_forward(sd, x, heads=None) -> Dict[str, torch.Tensor]:
    """-> {'hm','st','wh','ax','cr','reg'} raw head outputs at stride 4 (hm NOT yet sigmoid-ed) plus the intermediates."""
    out = lore_resnet18_features(sd, x)
    for head, _ in LORE_HEADS:
        if heads is None or head in heads:
            out[head] = lore_r18_head(sd, out["feat"], head)
    return out
