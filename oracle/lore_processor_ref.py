"""Plain-PyTorch fp32 restatement of the reference `LoreProcessModel` inference branch (TEST ORACLE).

Follows lore/lore_processor.py: LoreProcessModel.forward :465-514 (evaluation mode, wiz_stacking), Transformer :81-114,
Encoder :39-61 (positional encoder and final Norm are constructed but never applied), EncoderLayer :286-313 (pre-norm;
the last layer's extra attention map is discarded), MultiHeadAttention :172-226, attention :134-163, Norm :117-131
(UNBIASED std, eps added to the std), FeedForward :229-242, Decoder :64-78 (ends in ReLU), Stacker :342-396.
Pinned against the reference module by tests/golden/lore_processor_seed0.npz (oracle/gen_golden_lore.py).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F

HEADS = 8


def _t(sd, k):
    v = sd[k]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))


def _lin(x, sd, p):
    return F.linear(x, _t(sd, p + ".weight"), _t(sd, p + ".bias"))


def _norm(x, sd, p, eps=1e-6):
    mean = x.mean(-1, keepdim=True)
    std = x.std(-1, keepdim=True)  # unbiased
    return _t(sd, p + ".alpha") * (x - mean) / (std + eps) + _t(sd, p + ".bias")


def _mha(x, sd, p):
    n, d = x.shape
    dk = d // HEADS
    q = _lin(x, sd, p + ".q_linear").view(n, HEADS, dk).transpose(0, 1)
    k = _lin(x, sd, p + ".k_linear").view(n, HEADS, dk).transpose(0, 1)
    v = _lin(x, sd, p + ".v_linear").view(n, HEADS, dk).transpose(0, 1)
    s = torch.softmax(q @ k.transpose(-2, -1) / math.sqrt(dk), -1)
    return _lin((s @ v).transpose(0, 1).reshape(n, d), sd, p + ".out")


def _transformer(x, sd, p, n_layers):
    x = _lin(x, sd, p + ".linear")
    for L in range(n_layers):
        lp = f"{p}.encoder.layers.{L}"
        x = x + _mha(_norm(x, sd, lp + ".norm_1"), sd, lp + ".attn")
        x = x + _lin(F.relu(_lin(_norm(x, sd, lp + ".norm_2"), sd, lp + ".ff.linear_1")), sd, lp + ".ff.linear_2")
    return F.relu(_lin(F.relu(_lin(x, sd, p + ".decoder.linear.0")), sd, p + ".decoder.linear.2"))


@torch.no_grad()
def lore_processor_forward(sd, feat: torch.Tensor, layers: int = 4, stacking_layers: int = 4, dets=None):
    """feat fp32 [n,256] (one image) -> (logic_axis [n,4], stacked_axis [n,4]).  `dets` int64 [n,8] adds the 2-D
    position embeddings of the wireless / ptn configurations (:486-490); wtw passes None."""
    feat = feat.float()
    if dets is not None:
        xe, ye = _t(sd, "x_position_embeddings.weight"), _t(sd, "y_position_embeddings.weight")
        feat = feat + xe[dets[:, 0]] + ye[dets[:, 1]] + xe[dets[:, 2]] + ye[dets[:, 5]]
    logic = _transformer(feat, sd, "tsfm_axis", layers)
    emb = F.relu(_lin(F.relu(_lin(logic, sd, "stacker.logi_encoder.0")), sd, "stacker.logi_encoder.2"))
    stacked = _transformer(torch.cat((emb, feat), 1), sd, "stacker.tsfm", stacking_layers)
    return logic, stacked
