"""Plain-PyTorch fp32 restatement of the reference PULC classifier network (TEST ORACLE, see oracle/__init__.py).

Follows cls/cls_pp_lcnet.py: ConvBNLayer :73-100 (conv, BatchNorm, hardswish), DepthwiseSeparable :103-131, SEModule :134-160,
PPLCNet.forward :275-293 (conv1 -> blocks2..6 -> AdaptiveAvgPool2d(1) -> last_conv + hardswish -> flatten -> fc; dropout is
the identity in eval mode), with the task's stride list applied to the first block of blocks3..6 (:188-189).  Pinned against
the reference module itself by tests/golden/pulc_seed0.npz (oracle/gen_golden_pulc.py)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from pdf_table_b200.pplcnet_graph import net_config


def _t(sd, k):
    v = sd[k]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))


def _cbh(x, sd, p, stride=1, groups=1):
    w = _t(sd, p + ".conv.weight")
    x = F.conv2d(x, w, stride=stride, padding=(w.shape[-1] - 1) // 2, groups=groups)
    x = F.batch_norm(x, _t(sd, p + ".bn.running_mean"), _t(sd, p + ".bn.running_var"), _t(sd, p + ".bn.weight"), _t(sd, p + ".bn.bias"),
                     training=False, eps=1e-5)
    return F.hardswish(x)


def pplcnet_forward(sd, x: torch.Tensor, stride_list=(2, 2, 2, 2, 2)) -> torch.Tensor:
    """x fp32 [N,3,H,W] -> logits [N, class_num]."""
    with torch.no_grad():
        x = _cbh(x, sd, "conv1", stride=2)
        for name, cfg in net_config(stride_list).items():
            for i, (k, ci, co, s, se) in enumerate(cfg):
                p = f"{name}.{i}"
                x = _cbh(x, sd, p + ".dw_conv", stride=s, groups=ci)
                if se:
                    g = F.adaptive_avg_pool2d(x, 1)
                    g = F.relu(F.conv2d(g, _t(sd, p + ".se.conv1.weight"), _t(sd, p + ".se.conv1.bias")))
                    x = x * F.hardsigmoid(F.conv2d(g, _t(sd, p + ".se.conv2.weight"), _t(sd, p + ".se.conv2.bias")))
                x = _cbh(x, sd, p + ".pw_conv")
        x = F.hardswish(F.conv2d(F.adaptive_avg_pool2d(x, 1), _t(sd, "last_conv.weight")))
        return F.linear(x.flatten(1), _t(sd, "fc.weight"), _t(sd, "fc.bias"))
