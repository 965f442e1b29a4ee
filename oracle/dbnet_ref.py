"""Plain-PyTorch fp32 restatement of the reference DBNet-R18 forward (TEST ORACLE, see oracle/__init__.py).

Follows model/db_net/dbnet.py: ResNet.forward :322-334 (BasicBlock.forward :145-169, stride on conv1,
1x1 stride-s downsample :296-306), SegDetector.forward :618-640 (eval branch returns `binary`),
binarize head :533-539.  Pinned against the reference DBModel by tests/golden/dbnet_r18_seed0.npz
(oracle/gen_golden.py).
"""
from __future__ import annotations

from typing import Mapping

import numpy as np
import torch
import torch.nn.functional as F


def _t(sd, k):
    v = sd[k]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))


def _bn(x, sd, p, eps=1e-5):
    return F.batch_norm(x, _t(sd, p + ".running_mean"), _t(sd, p + ".running_var"), _t(sd, p + ".weight"),
                        _t(sd, p + ".bias"), training=False, eps=eps)


def _basic_block(x, sd, p, stride):
    out = F.relu(_bn(F.conv2d(x, _t(sd, p + ".conv1.weight"), stride=stride, padding=1), sd, p + ".bn1"))
    out = _bn(F.conv2d(out, _t(sd, p + ".conv2.weight"), padding=1), sd, p + ".bn2")
    if (p + ".downsample.0.weight") in sd:
        x = _bn(F.conv2d(x, _t(sd, p + ".downsample.0.weight"), stride=stride), sd, p + ".downsample.1")
    return F.relu(out + x)


@torch.no_grad()
def dbnet_r18_forward(sd: Mapping[str, np.ndarray], x: torch.Tensor, return_features: bool = False):
    """x: fp32 [N,3,H,W] -> probability map fp32 [N,1,H,W]."""
    x = x.to(_t(sd, "backbone.conv1.weight").dtype)  # fp32 oracle; the dtype of the weights when a half copy is timed on the GPU
    x = F.relu(_bn(F.conv2d(x, _t(sd, "backbone.conv1.weight"), stride=2, padding=3), sd, "backbone.bn1"))
    x = F.max_pool2d(x, 3, 2, 1)
    feats = []
    for L in range(1, 5):
        for B in range(2):
            x = _basic_block(x, sd, f"backbone.layer{L}.{B}", 2 if (L > 1 and B == 0) else 1)
        feats.append(x)
    c2, c3, c4, c5 = feats

    def conv(name, t, pad=0):
        b = sd.get(f"decoder.{name}.bias")
        return F.conv2d(t, _t(sd, f"decoder.{name}.weight"), None if b is None else _t(sd, f"decoder.{name}.bias"), padding=pad)

    in5, in4, in3, in2 = conv("in5", c5), conv("in4", c4), conv("in3", c3), conv("in2", c2)
    up = lambda t, s: F.interpolate(t, scale_factor=s, mode="nearest")
    out4 = up(in5, 2) + in4
    out3 = up(out4, 2) + in3
    out2 = up(out3, 2) + in2
    p5 = up(conv("out5.0", in5, 1), 8)
    p4 = up(conv("out4.0", out4, 1), 4)
    p3 = up(conv("out3.0", out3, 1), 2)
    p2 = conv("out2", out2, 1)
    fuse = torch.cat((p5, p4, p3, p2), 1)
    b = F.relu(_bn(conv("binarize.0", fuse, 1), sd, "decoder.binarize.1"))
    b = F.conv_transpose2d(b, _t(sd, "decoder.binarize.3.weight"), _t(sd, "decoder.binarize.3.bias"), stride=2)
    b = F.relu(_bn(b, sd, "decoder.binarize.4"))
    b = F.conv_transpose2d(b, _t(sd, "decoder.binarize.6.weight"), _t(sd, "decoder.binarize.6.bias"), stride=2)
    prob = torch.sigmoid(b)
    if return_features:
        return prob, dict(c2=c2, c3=c3, c4=c4, c5=c5, fuse=fuse)
    return prob
