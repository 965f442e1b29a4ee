"""CPU interpreter of the engine's graph programs (pdf_table_b200/picodet_graph.py -> csrc/graph_net.cu) -- TEST INFRASTRUCTURE.

Executes the lowered program (tensor table, op list, packed weights) with plain PyTorch fp32 ops, one op at a time with the
semantics of the CUDA executor (graph_net.cu: k_stem3x3s2, k_dwconv, the 1x1 conv_igemm plans, k_se_scale / k_se_apply,
k_up2, k_add, k_head_split), so that the LOWERING -- BatchNorm folding, weight packing, concatenations as channel slices, the
order of the ops -- can be checked on CPU against the oracle restatement of the reference modules (oracle/picodet_net_ref.py,
itself pinned against picodet/lcnet.py, csp_pan.py, pico_head.py).  Nothing here is product code.

`fp16_activations=True` rounds every op output to fp16 as the device buffers do (weights of the 1x1 convs are fp16 in the blob
already); with False the only difference from the oracle is the fp16 rounding of those weights.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

OP_STEM, OP_DW, OP_PW, OP_SE, OP_UP2, OP_ADD, OP_HEAD, OP_AVGPOOL, OP_UNFOLD3, OP_LN, OP_ATTN, OP_CTC, OP_CONV, OP_DECONV2, OP_DBHEAD = range(15)
ACT_NONE, ACT_RELU, ACT_HSWISH, ACT_SWISH = 0, 1, 4, 5


def _act(x: torch.Tensor, act: int) -> torch.Tensor:
    if act == ACT_NONE:
        return x
    if act == ACT_RELU:
        return F.relu(x)
    if act == ACT_HSWISH:
        return F.hardswish(x)
    if act == ACT_SWISH:
        return x * torch.sigmoid(x)
    raise ValueError(f"activation code {act} is not used by graph programs")


def run_program(blob: Dict[str, np.ndarray], x: torch.Tensor, fp16_activations: bool = False) -> Tuple[List[torch.Tensor], Dict[int, Tuple[torch.Tensor, torch.Tensor]]]:
    """blob: the tensor dict of build_picodet; x: fp32 [N,3,H,W] (pre-processed).  Returns (tensors NCHW fp32 by id,
    {level: (scores [N,HW,C], dfl [N,HW,R])})."""
    tensors, ops, meta = blob["graph.tensors"], blob["graph.ops"], blob["graph.meta"]
    num_classes, reg_bins = int(meta[0]), int(meta[1])
    n, _, hh, ww = x.shape
    rnd = (lambda t: t.half().float()) if fp16_activations else (lambda t: t)
    tens: List[torch.Tensor] = [x.float()]
    for row in tensors[1:]:
        c, dh, dw, ph, pw = (int(row[0]), int(row[1]), int(row[1]), 1, 1) if len(row) == 2 else (int(v) for v in row)
        tens.append(torch.zeros((n, c, 1 if ph == 0 else (-(-hh // dh)) // ph, 1 if pw == 0 else (-(-ww // dw)) // pw), dtype=torch.float32))
    heads: Dict[int, Tuple[torch.Tensor, torch.Tensor]] = {}

    def post_affine(t, wid):
        pa = blob.get(f"w{wid}.pa")
        return t if pa is None else t * float(pa[0]) + float(pa[1])

    def wt(wid: int, field: str) -> torch.Tensor:
        return torch.from_numpy(np.asarray(blob[f"w{wid}.{field}"]).astype(np.float32))

    for code, in_t, in_coff, in_c, out_t, out_coff, out_c, k, stride, act, w, aux in (tuple(int(v) for v in op) for op in ops):
        src = tens[in_t][:, in_coff:in_coff + in_c]
        if code == OP_STEM:  # sw [(ky*3 + kx)*3 + cin][cout]
            wk = wt(w, "sw").reshape(3, 3, 3, 16).permute(3, 2, 0, 1).contiguous()
            out = _act(F.conv2d(src, wk, wt(w, "sb"), stride=2, padding=1), act)
        elif code == OP_DW:  # dw [k*k][C]
            wk = wt(w, "dw").t().reshape(in_c, 1, k, k).contiguous()
            st = stride if stride < 256 else (stride & 255, stride >> 8)
            out = post_affine(_act(F.conv2d(src, wk, wt(w, "db"), stride=st, padding=k // 2, groups=in_c), act), w)
        elif code in (OP_PW, OP_HEAD, OP_CTC):  # w fp16 [Cout][Cin_pad], b fp32 padded
            wk = wt(w, "w")
            if code == OP_PW and k > 1:  # k pixels per GEMM row against diag(w, ..., w): every diagonal block is the layer's matrix
                full = wk[: k * out_c, : k * in_c]
                for j in range(k):
                    assert torch.equal(full[j * out_c:(j + 1) * out_c, j * in_c:(j + 1) * in_c], full[:out_c, :in_c]), "packed 1x1: unequal blocks"
                assert float(full.abs().sum()) == k * float(full[:out_c, :in_c].abs().sum()), "packed 1x1: off-diagonal weights"
                assert (n * src.shape[2] * src.shape[3]) % k == 0 and in_coff == 0 and out_coff == 0
                wk = full[:out_c]
            wk = wk[:, :in_c].reshape(-1, in_c, 1, 1)
            out = F.conv2d(src, wk, wt(w, "b")[:wk.shape[0]])
            if code == OP_PW and aux >= 0:
                out = out + tens[aux]
            out = post_affine(_act(out, act), w) if code == OP_PW else _act(out, act)
            if code == OP_CTC:  # k_softmax_rows over the first num_classes columns
                logits = out.permute(0, 2, 3, 1).reshape(n, -1, out.shape[1])[..., :num_classes]
                heads["logits"], heads["probs"] = logits, torch.softmax(logits, -1)
                continue
            if code == OP_HEAD:  # k_head_split: fp32 raw -> sigmoid class scores + raw DFL logits, pixel-major
                raw = out.permute(0, 2, 3, 1).reshape(n, -1, out.shape[1])
                heads[aux] = (torch.sigmoid(raw[..., :num_classes]), raw[..., num_classes:num_classes + reg_bins])
                continue
        elif code == OP_SE:
            avg = src.mean((2, 3))
            hid = F.relu(avg @ wt(w, "s1w").t() + wt(w, "s1b"))
            scale = torch.clamp((hid @ wt(w, "s2w").t() + wt(w, "s2b")) / 6.0 + 0.5, 0.0, 1.0)
            out = src * (scale + (1.0 if k == 2 else 0.0))[:, :, None, None]  # k = 2: RSELayer shortcut x + x * s
        elif code == OP_UP2:  # nearest, to the destination tensor's own size
            dst = tens[out_t]
            iy = torch.clamp(torch.arange(dst.shape[2]) * src.shape[2] // dst.shape[2], max=src.shape[2] - 1)
            ix = torch.clamp(torch.arange(dst.shape[3]) * src.shape[3] // dst.shape[3], max=src.shape[3] - 1)
            out = src[:, :, iy][:, :, :, ix]
            if aux >= 0:  # top-down FPN sum
                out = out + tens[aux]
        elif code == OP_ADD:
            out = tens[in_t] + tens[aux]
        elif code == OP_AVGPOOL:
            out = F.adaptive_avg_pool2d(src, 1) if k == 0 else F.avg_pool2d(src, [k & 255, k >> 8])
        elif code == OP_UNFOLD3:  # [N,C,1,T] -> [N,3C,1,T], channel = tap * C + c
            pad = F.pad(src, (1, 1))
            out = torch.cat([pad[..., t:t + src.shape[-1]] for t in range(3)], 1)
        elif code == OP_LN:
            out = F.layer_norm(src.permute(0, 2, 3, 1), (in_c,), wt(w, "lnw"), wt(w, "lnb"), eps=float(blob[f"w{w}.eps"][0])).permute(0, 3, 1, 2)
        elif code == OP_ATTN:  # qkv [N,3D,1,T]: q | k | v, head-major; the scale is folded into q
            d, t_len = out_c, src.shape[-1]
            qkv = src[:, :, 0].transpose(1, 2).reshape(n, t_len, 3, k, d // k).permute(2, 0, 3, 1, 4)
            attn = torch.softmax(qkv[0] @ qkv[1].transpose(-1, -2), -1)
            out = (attn @ qkv[2]).permute(0, 2, 1, 3).reshape(n, t_len, d).transpose(1, 2)[:, :, None]
        elif code == OP_CONV:  # dense k x k conv, stride 1: w fp16 [Cout][k*k*Cin_pad] (tap-major), b fp32 padded
            wk = wt(w, "w").reshape(out_c, k * k, -1)[:, :, :in_c].reshape(out_c, k, k, in_c).permute(0, 3, 1, 2).contiguous()
            out = _act(F.conv2d(src, wk, wt(w, "b")[:out_c], padding=k // 2), act)
        elif code == OP_DECONV2:  # ConvTranspose 2x2 s2 as a GEMM + pixel shuffle: w [(dy*2+dx)*Cout + co][Cin_pad]
            wk = wt(w, "w")[:, :in_c].reshape(2, 2, out_c, in_c).permute(3, 2, 0, 1).contiguous()  # -> [Cin, Cout, 2, 2]
            out = _act(F.conv_transpose2d(src, wk, wt(w, "b")[:out_c], stride=2), act)
        elif code == OP_DBHEAD:  # ConvTranspose 2x2 s2 C -> 1 + sigmoid -> fp32 probability map
            wk = wt(w, "hw").reshape(in_c, 1, 2, 2)
            heads["prob"] = torch.sigmoid(F.conv_transpose2d(src, wk, wt(w, "hb"), stride=2))
            continue
        else:
            raise ValueError(f"unknown opcode {code}")
        dst = tens[out_t]
        assert out.shape[1] == out_c and out.shape[2:] == dst.shape[2:], (code, tuple(out.shape), tuple(dst.shape), out_c)
        dst[:, out_coff:out_coff + out_c] = rnd(out)
    return tens, heads
