"""Plain-PyTorch fp32 restatement of the reference ConvNextViT recogniser (TEST ORACLE, see oracle/__init__.py).

Follows model/convnext_vit/modeling_convnext_vit.py:37-45 (RGB->gray, ConvNeXt, ViTForSTR),
modeling_convnext.py:28-131 (stages with (2,1) down-sampling; HF ConvNextEmbeddings / ConvNextLayer:
dw7x7 -> LN(eps 1e-6) -> Linear -> GELU -> Linear -> layer_scale -> +residual), modeling_vit.py:32-180
(patch projection 1x1, + position_embeddings[:,1:], 12 pre-LN layers eps 1e-12, final LN, 3x75 -> 201 token
stitch :135-139, classifier) and the pre/post-processors model/ocr_recognition/processor_ocr_recognition.py:
44-62, 73-115 (keep-ratio resize to 32 x <=804, zero pad, 3 chunks at x=0/252/504, /255) and :147-164
(argmax, collapse repeats, drop 0).  Pinned against the reference module by tests/golden/convnextvit_seed0.npz.
"""
from __future__ import annotations

from typing import List, Mapping

import numpy as np
import torch
import torch.nn.functional as F

DEPTHS = (3, 3, 8, 3)
DIMS = (96, 192, 256, 512)


def _t(sd, k):
    v = sd[k]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))


def to_torch(sd: Mapping, device="cpu", dtype=torch.float32):
    """state_dict as torch tensors on `device` in `dtype` (dtype=torch.float16 on CUDA reproduces the reference's
    own default inference precision, DeployUtils.model_eval utils/deploy_utils.py:226-240)."""
    return {k: _t(sd, k).to(device=device, dtype=dtype) for k in sd}


def _ln_cf(x, sd, p, eps):
    """LayerNorm over channels of an NCHW tensor (ConvNextLayerNorm channels_first)."""
    x = x.permute(0, 2, 3, 1)
    x = F.layer_norm(x, (x.shape[-1],), _t(sd, p + ".weight"), _t(sd, p + ".bias"), eps)
    return x.permute(0, 3, 1, 2)


@torch.no_grad()
def convnext_features(sd: Mapping, chunks: torch.Tensor) -> torch.Tensor:
    """chunks fp32 [B,3,32,300] in [0,1] -> last_hidden_state [B,512,1,75]."""
    x = chunks.to(_t(sd, "cnn_model.embeddings.patch_embeddings.weight").dtype)
    x = x[:, 0:1] * 0.2989 + x[:, 1:2] * 0.5870 + x[:, 2:3] * 0.1140
    p = "cnn_model.embeddings"
    x = F.conv2d(x, _t(sd, p + ".patch_embeddings.weight"), _t(sd, p + ".patch_embeddings.bias"), stride=4)
    x = _ln_cf(x, sd, p + ".layernorm", 1e-6)
    for s, depth in enumerate(DEPTHS):
        sp = f"cnn_model.encoder.stages.{s}"
        if s > 0:
            x = _ln_cf(x, sd, sp + ".downsampling_layer.0", 1e-6)
            x = F.conv2d(x, _t(sd, sp + ".downsampling_layer.1.weight"), _t(sd, sp + ".downsampling_layer.1.bias"), stride=(2, 1))
        for j in range(depth):
            lp = f"{sp}.layers.{j}"
            dim = x.shape[1]
            y = F.conv2d(x, _t(sd, lp + ".dwconv.weight"), _t(sd, lp + ".dwconv.bias"), padding=3, groups=dim)
            y = y.permute(0, 2, 3, 1)
            y = F.layer_norm(y, (dim,), _t(sd, lp + ".layernorm.weight"), _t(sd, lp + ".layernorm.bias"), 1e-6)
            y = F.linear(y, _t(sd, lp + ".pwconv1.weight"), _t(sd, lp + ".pwconv1.bias"))
            y = F.gelu(y)
            y = F.linear(y, _t(sd, lp + ".pwconv2.weight"), _t(sd, lp + ".pwconv2.bias"))
            y = _t(sd, lp + ".layer_scale_parameter") * y
            x = x + y.permute(0, 3, 1, 2)
    return x


@torch.no_grad()
def vit_tokens(sd: Mapping, feats: torch.Tensor, layers: int = 12, heads: int = 3) -> torch.Tensor:
    """[B,512,1,75] -> final-LN token features [B,75,192]."""
    v = "vitstr.vit"
    x = F.conv2d(feats, _t(sd, v + ".embeddings.patch_embeddings.projection.weight"),
                 _t(sd, v + ".embeddings.patch_embeddings.projection.bias"))
    x = x.flatten(2).transpose(1, 2)
    x = x + _t(sd, v + ".embeddings.position_embeddings")[:, 1:, :]
    B, T, D = x.shape
    hd = D // heads
    for L in range(layers):
        lp = f"{v}.encoder.layer.{L}"
        h = F.layer_norm(x, (D,), _t(sd, lp + ".layernorm_before.weight"), _t(sd, lp + ".layernorm_before.bias"), 1e-12)
        q, k, vv = (F.linear(h, _t(sd, f"{lp}.attention.attention.{n}.weight"), _t(sd, f"{lp}.attention.attention.{n}.bias"))
                    .view(B, T, heads, hd).transpose(1, 2) for n in ("query", "key", "value"))
        a = torch.softmax(q @ k.transpose(-1, -2) * (hd ** -0.5), dim=-1) @ vv
        a = a.transpose(1, 2).reshape(B, T, D)
        x = x + F.linear(a, _t(sd, lp + ".attention.output.dense.weight"), _t(sd, lp + ".attention.output.dense.bias"))
        h = F.layer_norm(x, (D,), _t(sd, lp + ".layernorm_after.weight"), _t(sd, lp + ".layernorm_after.bias"), 1e-12)
        h = F.gelu(F.linear(h, _t(sd, lp + ".intermediate.dense.weight"), _t(sd, lp + ".intermediate.dense.bias")))
        x = x + F.linear(h, _t(sd, lp + ".output.dense.weight"), _t(sd, lp + ".output.dense.bias"))
    return F.layer_norm(x, (D,), _t(sd, v + ".layernorm.weight"), _t(sd, v + ".layernorm.bias"), 1e-12)


def stitch(tok: torch.Tensor) -> torch.Tensor:
    """[3n,75,D] -> [n,201,D]  (modeling_vit.py:135-139)."""
    B, T, D = tok.shape
    ap = tok.view(B // 3, 3, T, D)
    return torch.cat((ap[:, 0, :69], ap[:, 1, 6:-6], ap[:, 2, 6:]), 1)


@torch.no_grad()
def convnextvit_forward(sd: Mapping, chunks: torch.Tensor, return_tokens: bool = False):
    """chunks fp32 [3n,3,32,300] -> logits [n,201,num_labels]."""
    tok = vit_tokens(sd, convnext_features(sd, chunks))
    x = stitch(tok)
    logits = F.linear(x, _t(sd, "vitstr.classifier.weight"), _t(sd, "vitstr.classifier.bias"))
    return (logits, tok) if return_tokens else logits


def keepratio_resize(img: np.ndarray, th: int = 32, tw: int = 804) -> np.ndarray:
    """processor_ocr_recognition.py:44-62 (cv2.resize bilinear, zero pad to th x tw)."""
    import cv2

    ratio = img.shape[1] / float(img.shape[0])
    cw = tw if ratio > float(tw) / th else int(th * ratio)
    img = cv2.resize(img, (cw, th))
    mask = np.zeros([th, tw, 3]).astype(np.uint8)
    mask[: img.shape[0], : img.shape[1], :] = img
    return mask


def preprocess(crops: List[np.ndarray]) -> torch.Tensor:
    """List of uint8 HWC crops -> fp32 [3n,3,32,300] (processor_ocr_recognition.py:73-115)."""
    out = []
    for c in crops:
        img = torch.FloatTensor(keepratio_resize(c))
        chunk = [img[:, 252 * i: 252 * i + 300] for i in range(3)]
        data = torch.cat(chunk, 0).view(3, 32, 300, 3) / 255.0
        out.append(data.permute(0, 3, 1, 2))
    return torch.cat(out, 0)


def greedy_ids(logits: torch.Tensor) -> List[np.ndarray]:
    """processor_ocr_recognition.py:147-162: softmax/argmax, keep p != last and p != 0."""
    preds = torch.argmax(F.softmax(logits, dim=-1), -1).cpu().numpy()
    out = []
    for row in preds:
        keep = np.ones(len(row), bool)
        keep[1:] = row[1:] != row[:-1]
        keep[0] = row[0] != 0
        keep &= row != 0
        out.append(row[keep].astype(np.int32))
    return out
