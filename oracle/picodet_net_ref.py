"""Plain-PyTorch fp32 restatement of the reference PicoDet forward (TEST ORACLE, see oracle/__init__.py).

Follows picodet/lcnet.py: ConvBNLayer :60-89, DepthwiseSeparable :92-127 (dw -> [SE] -> pw), SEModule :130-156, LCNet.forward
:239-258; picodet/csp_pan.py: DPModule :57-105, DarknetBottleneck :108-158, CSPLayer :161-207, Channel_T :210-230,
CSPPAN.forward :310-347 (top-down with nearest upsampling, bottom-up with stride-2 DPModules, the extra top level);
picodet/pico_head.py: PicoFeat.forward :151-167 (activation after EVERY dw / pw ConvNormLayer; the SE branch feeds only the
unused `reg` output when share_cls_reg), PicoHead.forward_eval :1108-1160 with export_post_process=False (:1130-1138):
per level sigmoid class scores [N, HW, C] and raw DFL logits [N, HW, 4*(reg_max+1)].
Pinned against the reference modules by tests/golden/picodet_net_seed0.npz (oracle/gen_golden_picodet.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from pdf_table_b200.synth import LCNET_CONFIG, PICO_HEAD_CONVS


def _t(sd, k):
    v = sd[k]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))


def _cba(x, sd, p, stride=1, groups=1, act=True, conv="conv", bn="bn"):
    w = _t(sd, f"{p}.{conv}.weight")
    x = F.conv2d(x, w, stride=stride, padding=(w.shape[-1] - 1) // 2, groups=groups)
    x = F.batch_norm(x, _t(sd, f"{p}.{bn}.running_mean"), _t(sd, f"{p}.{bn}.running_var"), _t(sd, f"{p}.{bn}.weight"),
                     _t(sd, f"{p}.{bn}.bias"), training=False, eps=1e-5)
    return F.hardswish(x) if act else x


def _dp(x, sd, p, stride=1):
    w = _t(sd, p + ".dwconv.weight")
    x = F.conv2d(x, w, stride=stride, padding=(w.shape[-1] - 1) // 2, groups=w.shape[0])
    x = F.hardswish(F.batch_norm(x, _t(sd, p + ".bn1.running_mean"), _t(sd, p + ".bn1.running_var"), _t(sd, p + ".bn1.weight"),
                                 _t(sd, p + ".bn1.bias"), training=False, eps=1e-5))
    x = F.conv2d(x, _t(sd, p + ".pwconv.weight"))
    return F.hardswish(F.batch_norm(x, _t(sd, p + ".bn2.running_mean"), _t(sd, p + ".bn2.running_var"), _t(sd, p + ".bn2.weight"),
                                    _t(sd, p + ".bn2.bias"), training=False, eps=1e-5))


def _csp(x, sd, p):
    short = _cba(x, sd, p + ".short_conv")
    main = _cba(x, sd, p + ".main_conv")
    main = _dp(_cba(main, sd, p + ".blocks.0.conv1"), sd, p + ".blocks.0.conv2")  # DarknetBottleneck, add_identity=False
    return _cba(torch.cat((main, short), 1), sd, p + ".final_conv")


def csppan_forward(sd, feats):
    t = [_cba(f, sd, f"conv_t.convs.{i}") for i, f in enumerate(feats)]
    inner = [t[-1]]
    for idx in range(len(t) - 1, 0, -1):
        up = F.interpolate(inner[0], size=t[idx - 1].shape[2:4], mode="nearest")
        inner.insert(0, _csp(torch.cat([up, t[idx - 1]], 1), sd, f"top_down_blocks.{len(t) - 1 - idx}"))
    outs = [inner[0]]
    for idx in range(len(t) - 1):
        down = _dp(outs[-1], sd, f"downsamples.{idx}", stride=2)
        outs.append(_csp(torch.cat([down, inner[idx + 1]], 1), sd, f"bottom_up_blocks.{idx}"))
    outs.append(_dp(t[-1], sd, "first_top_conv", stride=2) + _dp(outs[-1], sd, "second_top_conv", stride=2))
    return outs


def picohead_forward(sd, feats, num_classes):
    scores, dfl = [], []
    for lvl, x in enumerate(feats):
        for i in range(PICO_HEAD_CONVS):
            x = _cba(x, sd, f"conv_feat.cls_conv_dw{lvl}_{i}", groups=x.shape[1], bn="norm")
            x = _cba(x, sd, f"conv_feat.cls_conv_pw{lvl}_{i}", bn="norm")
        y = F.conv2d(x, _t(sd, f"head_cls{lvl}.weight"), _t(sd, f"head_cls{lvl}.bias"))
        n = y.shape[0]
        scores.append(torch.sigmoid(y[:, :num_classes]).reshape(n, num_classes, -1).permute(0, 2, 1))
        dfl.append(y[:, num_classes:].reshape(n, y.shape[1] - num_classes, -1).permute(0, 2, 1))
    return scores, dfl


@torch.no_grad()
def picodet_forward(backbone_sd, neck_sd, head_sd, x: torch.Tensor, num_classes: int = 5, return_features: bool = False):
    """x fp32 [N,3,H,W] (pre-processed) -> (scores[4] fp32 [N,HW_l,C], dfl[4] fp32 [N,HW_l,32])."""
    c = lcnet_all(backbone_sd, x.to(_t(backbone_sd, "conv1.conv.weight").dtype))
    neck = csppan_forward(neck_sd, c)
    s, d = picohead_forward(head_sd, neck, num_classes)
    if return_features:
        return s, d, {"c3": c[0], "c4": c[1], "c5": c[2], "p3": neck[0], "p4": neck[1], "p5": neck[2], "p6": neck[3]}
    return s, d


def lcnet_all(sd, x):
    """LCNet.forward :239-258 with feature_maps [3, 4, 5]: the outputs of blocks4, blocks5, blocks6."""
    x = _cba(x, sd, "conv1", stride=2)
    outs = {}
    for name, cfg in LCNET_CONFIG.items():
        for i, (k, cin, cout, s, se) in enumerate(cfg):
            p = f"{name}.{i}"
            x = _cba(x, sd, p + ".dw_conv", stride=s, groups=cin)
            if se:
                a = F.adaptive_avg_pool2d(x, 1)
                a = F.relu(F.conv2d(a, _t(sd, p + ".se.conv1.weight"), _t(sd, p + ".se.conv1.bias")))
                a = F.hardsigmoid(F.conv2d(a, _t(sd, p + ".se.conv2.weight"), _t(sd, p + ".se.conv2.bias")))
                x = x * a
            x = _cba(x, sd, p + ".pw_conv")
        outs[name] = x
    return [outs["blocks4"], outs["blocks5"], outs["blocks6"]]
