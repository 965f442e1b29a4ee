"""Plain-PyTorch fp32 restatement of the PP-OCRv4 mobile text detector (PPLCNetV3-0.75 backbone -> RSE-FPN -> DBHead),
SURVEY.md row a2 (TEST ORACLE, see oracle/__init__.py).

PARITY UNPINNED against the reference's own inference: the reference runs this network as an ONNX file downloaded from the
hub (`cycloneboy/{ch,en,...}_PP-OCRv4_det_infer`, ocr_pdf/ocr_table_model_config.py:134-147, executed at
ocr_pdf/ocr_detection_task.py:98-107); neither the graph, nor weights, nor onnx / onnxruntime exist in this image, and the
reference repository contains no PP-OCRv4 architecture code (its in-tree detector is DBNet-R18, oracle/dbnet_ref.py).  What is
restated here is the PUBLISHED architecture the ONNX was exported from (PaddleOCR release 2.7, Apache-2.0; SURVEY.md 8c names
the same structure: PPLCNetV3-0.75 -> RSE-FPN(96) -> DBHead(k=50), rep branches fused at export):

  * ppocr/modeling/backbones/rec_lcnetv3.py -- PPLCNetV3(scale=0.75, det=True) in its DEPLOY form: conv1 3x3 s2 + BN (no
    activation); LCNetV3Block = depthwise rep layer -> [SE] -> pointwise rep layer; rep layer = conv -> LearnableAffineBlock
    -> (hardswish -> LearnableAffineBlock) unless the layer's stride is 2; four taps (after blocks3..6, strides 4 / 8 / 16 /
    32) each through a 1x1 conv with bias to int(mv_c * 0.75) = 12 / 18 / 42 / 360 channels.  Channels
    (make_divisible(c * 0.75, 16)): 16, 32, 48, 96, 192, 384.
  * ppocr/modeling/necks/db_fpn.py RSEFPN(out_channels=96, shortcut=True): ins_conv = RSELayer(c_i, 96, k=1), top-down
    nearest x2 sums, inp_conv = RSELayer(96, 24, k=3), nearest x8 / x4 / x2, concat [p5, p4, p3, p2]; RSELayer = conv (no
    bias) -> x + SEModule(x); SEModule (det_mobilenet_v3.py): avg-pool -> conv C/4 + relu -> conv C ->
    hardsigmoid(slope 0.2, offset 0.5) -> x * s.
  * ppocr/modeling/heads/det_db_head.py DBHead.binarize (inference returns the shrink map only): conv3x3 96 -> 24 (no bias)
    + BN + relu -> ConvTranspose 2x2 s2 24 -> 24 + BN + relu -> ConvTranspose 2x2 s2 24 -> 1 -> sigmoid.

Input fp32 [B, 3, H, W] (H, W multiples of 32: DetResizeForTest, a1) -> probability map fp32 [B, 1, H, W] (the tensor
DBPostProcess, a3, consumes).  State-dict keys follow the Paddle module tree with torch conventions.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# k, in_c, out_c, stride, use_se (rec_lcnetv3.py NET_CONFIG_det)
NET_CONFIG_DET = {
    "blocks2": [[3, 16, 32, 1, False]],
    "blocks3": [[3, 32, 64, 2, False], [3, 64, 64, 1, False]],
    "blocks4": [[3, 64, 128, 2, False], [3, 128, 128, 1, False]],
    "blocks5": [[3, 128, 256, 2, False], [5, 256, 256, 1, False], [5, 256, 256, 1, False], [5, 256, 256, 1, False],
                [5, 256, 256, 1, False]],
    "blocks6": [[5, 256, 512, 2, True], [5, 512, 512, 1, True], [5, 512, 512, 1, False], [5, 512, 512, 1, False]],
}
SCALE = 0.75
MV_C = (16, 24, 56, 480)  # tap widths before the scale: int(c * 0.75) = 12, 18, 42, 360
FPN_C = 96


def make_divisible(v, divisor=16, min_value=None):
    if min_value is None:
        min_value = divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


def ch(c: int) -> int:
    return make_divisible(c * SCALE)


def tap_channels():
    return [int(c * SCALE) for c in MV_C]


def _t(sd, k):
    v = sd[k]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))


def _bn(x, sd, p):
    return F.batch_norm(x, _t(sd, p + ".running_mean"), _t(sd, p + ".running_var"), _t(sd, p + ".weight"), _t(sd, p + ".bias"),
                        training=False, eps=1e-5)


def _lab(x, sd, p):
    return _t(sd, p + ".scale") * x + _t(sd, p + ".bias")


def _rep(x, sd, p, stride, groups):
    """Deploy-form LearnableRepLayer.forward: lab(reparam_conv(x)), then Act (hardswish + lab) unless stride == 2."""
    w = _t(sd, p + ".reparam_conv.weight")
    k = w.shape[-1]
    x = F.conv2d(x, w, _t(sd, p + ".reparam_conv.bias"), stride=stride, padding=(k - 1) // 2, groups=groups)
    x = _lab(x, sd, p + ".lab")
    if stride != 2:
        x = _lab(F.hardswish(x), sd, p + ".act.lab")
    return x


def _se_backbone(x, sd, p):
    s = F.adaptive_avg_pool2d(x, 1)
    s = F.relu(F.conv2d(s, _t(sd, p + ".conv1.weight"), _t(sd, p + ".conv1.bias")))
    s = F.hardsigmoid(F.conv2d(s, _t(sd, p + ".conv2.weight"), _t(sd, p + ".conv2.bias")))
    return x * s


def backbone_forward(sd, x):
    """-> the four tap tensors [B, 12 / 18 / 42 / 360, H/4 .. H/32, W/4 .. W/32]."""
    x = _bn(F.conv2d(x, _t(sd, "backbone.conv1.conv.weight"), stride=2, padding=1), sd, "backbone.conv1.bn")
    outs = []
    for name, cfg in NET_CONFIG_DET.items():
        for i, (k, cin, cout, s, se) in enumerate(cfg):
            p = f"backbone.{name}.{i}"
            x = _rep(x, sd, p + ".dw_conv", s, ch(cin))
            if se:
                x = _se_backbone(x, sd, p + ".se")
            x = _rep(x, sd, p + ".pw_conv", 1, 1)
        if name != "blocks2":
            outs.append(x)
    return [F.conv2d(o, _t(sd, f"backbone.layer_list.{i}.weight"), _t(sd, f"backbone.layer_list.{i}.bias")) for i, o in enumerate(outs)]


def _rse(x, sd, p, k):
    """RSELayer: conv (no bias) -> x + x * hardsigmoid_{0.2, 0.5}(conv2(relu(conv1(avgpool(x)))))."""
    x = F.conv2d(x, _t(sd, p + ".in_conv.weight"), padding=k // 2)
    s = F.adaptive_avg_pool2d(x, 1)
    s = F.relu(F.conv2d(s, _t(sd, p + ".se_block.conv1.weight"), _t(sd, p + ".se_block.conv1.bias")))
    s = torch.clamp(0.2 * F.conv2d(s, _t(sd, p + ".se_block.conv2.weight"), _t(sd, p + ".se_block.conv2.bias")) + 0.5, 0.0, 1.0)
    return x + x * s


def neck_forward(sd, taps):
    c2, c3, c4, c5 = taps
    in5, in4 = _rse(c5, sd, "neck.ins_conv.3", 1), _rse(c4, sd, "neck.ins_conv.2", 1)
    in3, in2 = _rse(c3, sd, "neck.ins_conv.1", 1), _rse(c2, sd, "neck.ins_conv.0", 1)
    out4 = in4 + F.interpolate(in5, scale_factor=2, mode="nearest")
    out3 = in3 + F.interpolate(out4, scale_factor=2, mode="nearest")
    out2 = in2 + F.interpolate(out3, scale_factor=2, mode="nearest")
    p5, p4 = _rse(in5, sd, "neck.inp_conv.3", 3), _rse(out4, sd, "neck.inp_conv.2", 3)
    p3, p2 = _rse(out3, sd, "neck.inp_conv.1", 3), _rse(out2, sd, "neck.inp_conv.0", 3)
    p5 = F.interpolate(p5, scale_factor=8, mode="nearest")
    p4 = F.interpolate(p4, scale_factor=4, mode="nearest")
    p3 = F.interpolate(p3, scale_factor=2, mode="nearest")
    return torch.cat([p5, p4, p3, p2], 1)


def head_forward(sd, x):
    p = "head.binarize"
    x = F.relu(_bn(F.conv2d(x, _t(sd, p + ".conv1.weight"), padding=1), sd, p + ".conv_bn1"))
    x = F.relu(_bn(F.conv_transpose2d(x, _t(sd, p + ".conv2.weight"), _t(sd, p + ".conv2.bias"), stride=2), sd, p + ".conv_bn2"))
    return torch.sigmoid(F.conv_transpose2d(x, _t(sd, p + ".conv3.weight"), _t(sd, p + ".conv3.bias"), stride=2))


def pp_det_forward(sd, x: torch.Tensor, return_fuse: bool = False):
    """x fp32 [B,3,H,W] (normalised as PPOcrDetectionPreprocessor does) -> probability map [B,1,H,W]."""
    with torch.no_grad():
        fuse = neck_forward(sd, backbone_forward(sd, x))
        prob = head_forward(sd, fuse)
    return (prob, fuse) if return_fuse else prob
