"""Plain-PyTorch fp32 restatement of the reference CRNN recogniser (TEST ORACLE, see oracle/__init__.py).

Follows crnn/modeling_crnn.py: BidirectionalLSTM.forward (:21-33) and CRNN.forward (:90-113) -- RGB -> gray with the
reference's coefficients, conv0..conv4 (+ BatchNorm in eval mode + ReLU) with max-pools (2,2) (2,2) (2,1) (2,1), the sequence
[w, b, c] through two bidirectional LSTMs (hidden 256) each followed by a Linear over the concatenated directions, the
bias-free classifier, output [b, w / 4, labels].  Pinned against the reference module itself by tests/golden/crnn_seed0.npz
(oracle/gen_golden_crnn.py runs `CRNN` from /root/reference on the same seeded weights; tests/test_oracle_cpu.py).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def _t(sd, k):
    v = sd[k]
    return v if isinstance(v, torch.Tensor) else torch.from_numpy(np.asarray(v))


def _cbr(x, sd, conv, bn, stride=1, padding=1):
    x = F.conv2d(x, _t(sd, conv + ".weight"), _t(sd, conv + ".bias"), stride=stride, padding=padding)
    x = F.batch_norm(x, _t(sd, bn + ".running_mean"), _t(sd, bn + ".running_var"), _t(sd, bn + ".weight"), _t(sd, bn + ".bias"), training=False, eps=1e-5)
    return F.relu(x)


def _lstm_dir(x, w_ih, w_hh, b_ih, b_hh, reverse):
    """One direction of nn.LSTM over x [T, B, In] (gate order i, f, g, o); returns [T, B, H]."""
    T, B, _ = x.shape
    H = w_hh.shape[1]
    h = torch.zeros(B, H)
    c = torch.zeros(B, H)
    out = [None] * T
    for t in (range(T - 1, -1, -1) if reverse else range(T)):
        g = x[t] @ w_ih.t() + b_ih + h @ w_hh.t() + b_hh
        i, f, gg, o = g[:, :H], g[:, H:2 * H], g[:, 2 * H:3 * H], g[:, 3 * H:]
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(gg)
        h = torch.sigmoid(o) * torch.tanh(c)
        out[t] = h
    return torch.stack(out)


def _bilstm(x, sd, p):
    r = p + ".rnn."
    fwd = _lstm_dir(x, _t(sd, r + "weight_ih_l0"), _t(sd, r + "weight_hh_l0"), _t(sd, r + "bias_ih_l0"), _t(sd, r + "bias_hh_l0"), False)
    bwd = _lstm_dir(x, _t(sd, r + "weight_ih_l0_reverse"), _t(sd, r + "weight_hh_l0_reverse"), _t(sd, r + "bias_ih_l0_reverse"),
                    _t(sd, r + "bias_hh_l0_reverse"), True)
    rec = torch.cat([fwd, bwd], 2)
    T, B, hh = rec.shape
    return F.linear(rec.reshape(T * B, hh), _t(sd, p + ".embedding.weight"), _t(sd, p + ".embedding.bias")).reshape(T, B, -1)


def crnn_features(sd, x: torch.Tensor) -> torch.Tensor:
    """x fp32 [B,3,32,W] in [0,1] -> conv features as the sequence [W/4, B, 512]."""
    x = x[:, 0:1] * 0.2989 + x[:, 1:2] * 0.5870 + x[:, 2:3] * 0.1140
    x = F.max_pool2d(_cbr(x, sd, "conv0.0", "conv0.1"), 2, 2)
    x = F.max_pool2d(_cbr(x, sd, "conv1.0", "conv1.1"), 2, 2)
    x = _cbr(_cbr(x, sd, "conv2.0", "conv2.1"), sd, "conv2.3", "conv2.4")
    x = F.max_pool2d(x, (2, 1), (2, 1))
    x = _cbr(_cbr(x, sd, "conv3.0", "conv3.1"), sd, "conv3.3", "conv3.4")
    x = F.max_pool2d(x, (2, 1), (2, 1))
    x = _cbr(x, sd, "conv4.0", "conv4.1", stride=(2, 1), padding=0)
    assert x.shape[2] == 1, "the height of conv must be 1"
    return x.squeeze(2).permute(2, 0, 1)


def crnn_forward(sd, x: torch.Tensor) -> torch.Tensor:
    """x fp32 [B,3,32,W] in [0,1] -> logits [B, W/4, labels]."""
    with torch.no_grad():
        seq = crnn_features(sd, x)
        seq = _bilstm(_bilstm(seq, sd, "rnn.0"), sd, "rnn.1")
        return F.linear(seq, _t(sd, "cls.weight")).permute(1, 0, 2)
