"""Plain-PyTorch fp32 restatement of the PP-OCRv4 text recogniser ("SVTR-LCNet": PPLCNetV3-0.95 backbone -> SVTR neck -> CTC
head), SURVEY.md row a5 (TEST ORACLE, see oracle/__init__.py).

PARITY UNPINNED against the reference's own inference: the reference runs this network as an ONNX file downloaded from the
hub (`cycloneboy/{ch,en,...}_PP-OCRv4_rec_infer`, ocr_pdf/ocr_table_model_config.py:166-204, executed at
ocr_pdf/ocr_recognition_task.py:90-99); neither the graph, nor weights, nor onnx / onnxruntime exist in this image, and the
reference repository contains no PP-OCRv4 architecture code.  What is restated here is the PUBLISHED architecture the ONNX was
exported from (PaddleOCR release 2.7, Apache-2.0; SURVEY.md 8c names the same structure):

  * ppocr/modeling/backbones/rec_lcnetv3.py -- PPLCNetV3(scale=0.95, det=False) in its DEPLOY form (every LearnableRepLayer
    re-parameterised to one conv + bias, as `export_model` does): conv1 3x3 s2 (no activation); LCNetV3Block = depthwise rep
    layer (k 3 / 5, strides 1, (2,1), (1,2)) -> [SE] -> pointwise rep layer; rep layer = conv -> LearnableAffineBlock (scalar
    scale, bias) -> hardswish -> LearnableAffineBlock (the activation is skipped only for an INT stride of 2, which the rec
    config never uses); eval tail avg_pool2d(x, [3, 2]).  Channels (make_divisible(c * 0.95, 16)): 16, 32, 64, 128, 240, 480.
  * ppocr/modeling/necks/rnn.py EncoderWithSVTR(dims=120, depth=2, hidden_dims=120, kernel_size=[1, 3], use_guide=True) with
    ppocr/modeling/backbones/rec_svtrnet.py Block(mixer="Global", num_heads=8, mlp_ratio=2, act=Swish, eps 1e-5,
    prenorm=False -> x + mixer(norm1(x)), x + mlp(norm2(x))); final LayerNorm eps 1e-6.
  * ppocr/modeling/heads/rec_ctc_head.py CTCHead: Linear(120, n_class) + softmax at inference.

Input fp32 [B, 3, 48, W] (PPOcrRecPreProcessor's output, a4) -> probabilities fp32 [B, W // 8, n_class] (the tensor
CTCLabelDecode, a6, consumes).  State-dict keys follow the Paddle module tree with torch conventions (Linear weight
[out, in]; BatchNorm weight / bias / running_mean / running_var).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

# k, in_c, out_c, stride, use_se (rec_lcnetv3.py NET_CONFIG_rec)
NET_CONFIG_REC = {
    "blocks2": [[3, 16, 32, 1, False]],
    "blocks3": [[3, 32, 64, 1, False], [3, 64, 64, 1, False]],
    "blocks4": [[3, 64, 128, (2, 1), False], [3, 128, 128, 1, False]],
    "blocks5": [[3, 128, 256, (1, 2), False], [5, 256, 256, 1, False], [5, 256, 256, 1, False], [5, 256, 256, 1, False],
                [5, 256, 256, 1, False]],
    "blocks6": [[5, 256, 512, (2, 1), True], [5, 512, 512, 1, True], [5, 512, 512, (2, 1), False], [5, 512, 512, 1, False]],
}
SCALE = 0.95
NECK_DIMS, NECK_HIDDEN, NECK_DEPTH, NECK_HEADS, NECK_MLP = 120, 120, 2, 8, 2.0


def make_divisible(v, divisor=16, min_value=None):
    if min_value is None:
        min_value = divisor
    new_v = max(min_value, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


def ch(c: int) -> int:
    return make_divisible(c * SCALE)


def _t(sd, k):
    v = sd[k]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))


def _bn(x, sd, p):
    return F.batch_norm(x, _t(sd, p + ".running_mean"), _t(sd, p + ".running_var"), _t(sd, p + ".weight"), _t(sd, p + ".bias"),
                        training=False, eps=1e-5)


def _lab(x, sd, p):
    return _t(sd, p + ".scale") * x + _t(sd, p + ".bias")


def _rep(x, sd, p, stride, groups):
    """Deploy-form LearnableRepLayer.forward: lab(reparam_conv(x)), then Act (hardswish + lab) unless stride == 2 (an int)."""
    w = _t(sd, p + ".reparam_conv.weight")
    k = w.shape[-1]
    x = F.conv2d(x, w, _t(sd, p + ".reparam_conv.bias"), stride=stride, padding=(k - 1) // 2, groups=groups)
    x = _lab(x, sd, p + ".lab")
    if stride != 2:
        x = _lab(F.hardswish(x), sd, p + ".act.lab")
    return x


def _se(x, sd, p):
    s = F.adaptive_avg_pool2d(x, 1)
    s = F.relu(F.conv2d(s, _t(sd, p + ".conv1.weight"), _t(sd, p + ".conv1.bias")))
    s = F.hardsigmoid(F.conv2d(s, _t(sd, p + ".conv2.weight"), _t(sd, p + ".conv2.bias")))
    return x * s


def backbone_forward(sd, x, return_features: bool = False):
    feats = {}
    x = _bn(F.conv2d(x, _t(sd, "backbone.conv1.conv.weight"), stride=2, padding=1), sd, "backbone.conv1.bn")
    for name, cfg in NET_CONFIG_REC.items():
        for i, (k, cin, cout, s, se) in enumerate(cfg):
            p = f"backbone.{name}.{i}"
            x = _rep(x, sd, p + ".dw_conv", s, ch(cin))
            if se:
                x = _se(x, sd, p + ".se")
            x = _rep(x, sd, p + ".pw_conv", 1, 1)
        feats[name] = x
    x = F.avg_pool2d(x, [3, 2])
    return (x, feats) if return_features else x


def _cbs(x, sd, p, pad):
    """rnn.py ConvBNLayer: conv (no bias) + BN + Swish."""
    x = _bn(F.conv2d(x, _t(sd, p + ".conv.weight"), padding=pad), sd, p + ".norm")
    return x * torch.sigmoid(x)


def _svtr_block(x, sd, p):
    b, n, c = x.shape
    h = F.layer_norm(x, (c,), _t(sd, p + ".norm1.weight"), _t(sd, p + ".norm1.bias"), eps=1e-5)
    qkv = F.linear(h, _t(sd, p + ".mixer.qkv.weight"), _t(sd, p + ".mixer.qkv.bias")).reshape(b, n, 3, NECK_HEADS, c // NECK_HEADS).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (c // NECK_HEADS) ** -0.5, qkv[1], qkv[2]
    attn = F.softmax(q @ k.transpose(-1, -2), dim=-1)
    h = (attn @ v).permute(0, 2, 1, 3).reshape(b, n, c)
    x = x + F.linear(h, _t(sd, p + ".mixer.proj.weight"), _t(sd, p + ".mixer.proj.bias"))
    h = F.layer_norm(x, (c,), _t(sd, p + ".norm2.weight"), _t(sd, p + ".norm2.bias"), eps=1e-5)
    h = F.linear(h, _t(sd, p + ".mlp.fc1.weight"), _t(sd, p + ".mlp.fc1.bias"))
    h = h * torch.sigmoid(h)
    return x + F.linear(h, _t(sd, p + ".mlp.fc2.weight"), _t(sd, p + ".mlp.fc2.bias"))


def neck_forward(sd, x):
    """EncoderWithSVTR.forward + Im2Seq: [B, 480, 1, T] -> [B, T, 120]."""
    p = "head.ctc_encoder.encoder"
    h = x
    z = _cbs(x, sd, p + ".conv1", (0, 1))
    z = _cbs(z, sd, p + ".conv2", 0)
    b, c, hh, ww = z.shape
    z = z.flatten(2).transpose(1, 2)
    for i in range(NECK_DEPTH):
        z = _svtr_block(z, sd, f"{p}.svtr_block.{i}")
    z = F.layer_norm(z, (c,), _t(sd, p + ".norm.weight"), _t(sd, p + ".norm.bias"), eps=1e-6)
    z = z.reshape(b, hh, ww, c).permute(0, 3, 1, 2)
    z = _cbs(z, sd, p + ".conv3", 0)
    z = torch.cat((h, z), 1)
    z = _cbs(_cbs(z, sd, p + ".conv4", (0, 1)), sd, p + ".conv1x1", 0)
    return z.squeeze(2).transpose(1, 2)


def pp_rec_forward(sd, x: torch.Tensor, return_logits: bool = False):
    """x fp32 [B,3,48,W] -> softmax probabilities [B, W // 8, n_class] (and the logits before the softmax)."""
    with torch.no_grad():
        seq = neck_forward(sd, backbone_forward(sd, x))
        logits = F.linear(seq, _t(sd, "head.ctc_head.fc.weight"), _t(sd, "head.ctc_head.fc.bias"))
        probs = F.softmax(logits, dim=2)
    return (probs, logits) if return_logits else probs
