"""PicoDet (LCNet-x1.0 + CSP-PAN + PicoHead) as a graph program for the engine's generic executor (csrc/graph_net.cu).

The reference assembles the detector from config dictionaries (`PicoDet(backbone_config, neck_config, head_config)`,
picodet/modeling_picodet.py:32-36) and ships only the ONNX export; the in-tree modules are picodet/lcnet.py:159-263,
picodet/csp_pan.py:233-360 and picodet/pico_head.py:37-167, 972-1160.  Instead of a hand-written C++ plan per network, this
module lowers the architecture + a reference state_dict to

  * a tensor table  [n_tensors][2]  int32: channels, down-sampling factor (spatial size = ceil(H / down) x ceil(W / down));
  * a program       [n_ops][12]     int32: opcode, in_t, in_coff, in_c, out_t, out_coff, out_c, k, stride, act, w_id, aux;
  * weights `w{id}.*`: BatchNorm folded, 1x1 convs packed K-major fp16 for conv_igemm_tcgen05, depthwise kernels fp32
    [k*k][C], SE / stem weights fp32.

Concatenations are channel slices of pre-planned buffers (in_coff / out_coff), exactly like the DLA roots of the Lore
detector; nothing is copied.  The PicoSE branch of PicoFeat is not lowered: with share_cls_reg its output only feeds the
`reg` tensor that PicoHead.forward_eval never reads (pico_head.py:1117-1124).
"""
from __future__ import annotations

from typing import Dict, List, Mapping, Tuple

import numpy as np

from . import weights as W
from .synth import LCNET_CONFIG, PICO_HEAD_CONVS, PICO_NECK_CH

OP_STEM, OP_DW, OP_PW, OP_SE, OP_UP2, OP_ADD, OP_HEAD = range(7)
ACT_NONE, ACT_HSWISH = 0, 4  # dv_act codes


class _Builder:
    def __init__(self):
        self.tensors: List[Tuple[int, int]] = []
        self.ops: List[List[int]] = []
        self.blob: Dict[str, np.ndarray] = {}
        self.nw = 0

    def tensor(self, c: int, down: int) -> int:
        self.tensors.append((c, down))
        return len(self.tensors) - 1

    def op(self, code, in_t, out_t, in_coff=0, in_c=None, out_coff=0, out_c=None, k=1, stride=1, act=ACT_NONE, w=-1, aux=-1):
        in_c = self.tensors[in_t][0] if in_c is None else in_c
        out_c = self.tensors[out_t][0] if out_c is None else out_c
        self.ops.append([code, in_t, in_coff, in_c, out_t, out_coff, out_c, k, stride, act, w, aux])

    def weight(self, **arrays) -> int:
        for k, v in arrays.items():
            self.blob[f"w{self.nw}.{k}"] = np.ascontiguousarray(v)
        self.nw += 1
        return self.nw - 1


def _f(sd, k):
    return W._np(sd[k]).astype(np.float32)


def _bn(sd, p):
    return {k: W._np(sd[f"{p}.{k}"]) for k in ("weight", "bias", "running_mean", "running_var")}


def _pw(b: _Builder, sd, conv_key, bn_prefix, bias_key=None, pack: int = 1):
    """1x1 conv + folded BN.  pack > 1: the block-diagonal weight of `pack` consecutive pixels per GEMM row (OP_PW with k = pack;
    see pp_rec_graph.pw_pack_factor for why the 16- / 32-channel layers run that way)."""
    bias_in, bn = None if bias_key is None else _f(sd, bias_key), None if bn_prefix is None else _bn(sd, bn_prefix)
    if pack == 1:
        w, bias = W.pack_conv(_f(sd, conv_key), bias_in, bn)
        return b.weight(w=w, b=bias)
    cout, cin = (int(v) for v in _f(sd, conv_key).shape[:2])
    w32, b32, _ = W.conv_matrix_f32(_f(sd, conv_key), bias_in, bn)
    wd = np.zeros((pack * cout, pack * cin), np.float32)
    for j in range(pack):
        wd[j * cout:(j + 1) * cout, j * cin:(j + 1) * cin] = w32[:, :cin]
    return b.weight(w=wd.astype(np.float16), b=W.pad_bias(np.tile(b32[:cout], pack)))


def pw_pack(cin: int, cout: int) -> int:
    return 64 // cin if (cin <= 32 and cin % 16 == 0 and cout * (64 // cin) <= 256) else 1


def _dw(b: _Builder, sd, conv_key, bn_prefix):
    w = _f(sd, conv_key)  # [C,1,k,k]
    scale, shift = W.bn_affine(_bn(sd, bn_prefix), w.shape[0])
    k = w.shape[-1]
    return b.weight(dw=(w[:, 0] * scale[:, None, None]).transpose(1, 2, 0).reshape(k * k, -1).astype(np.float32), db=shift.astype(np.float32)), k


def build_picodet(backbone: Mapping, neck: Mapping, head: Mapping, num_classes: int):
    """-> (blob tensors dict, meta) for weights.write_blob; see the module docstring for the program format."""
    b = _Builder()
    # ---- LCNet
    img = b.tensor(3, 1)
    w = _f(backbone, "conv1.conv.weight")  # [16,3,3,3]
    scale, shift = W.bn_affine(_bn(backbone, "conv1.bn"), 16)
    x = b.tensor(16, 2)
    b.op(OP_STEM, img, x, k=3, stride=2, act=ACT_HSWISH,
         w=b.weight(sw=(w * scale[:, None, None, None]).transpose(2, 3, 1, 0).reshape(27, 16).astype(np.float32), sb=shift.astype(np.float32)))
    down = 2
    feats = {}
    # C4 / C5 are written straight into their consumers' concatenation buffers by conv_t below, C3..C5 themselves are dense
    for name, cfg in LCNET_CONFIG.items():
        for i, (k, cin, cout, s, se) in enumerate(cfg):
            p = f"{name}.{i}"
            down *= s
            t = b.tensor(cin, down)
            wid, kk = _dw(b, backbone, p + ".dw_conv.conv.weight", p + ".dw_conv.bn")
            b.op(OP_DW, x, t, k=kk, stride=s, act=ACT_HSWISH, w=wid)
            if se:
                t2 = b.tensor(cin, down)
                b.op(OP_SE, t, t2, w=b.weight(s1w=_f(backbone, p + ".se.conv1.weight").reshape(cin // 4, cin), s1b=_f(backbone, p + ".se.conv1.bias"),
                                             s2w=_f(backbone, p + ".se.conv2.weight").reshape(cin, cin // 4), s2b=_f(backbone, p + ".se.conv2.bias")))
                t = t2
            x = b.tensor(cout, down)
            pk = pw_pack(cin, cout)
            b.op(OP_PW, t, x, k=pk, act=ACT_HSWISH, w=_pw(b, backbone, p + ".pw_conv.conv.weight", p + ".pw_conv.bn", pack=pk))
        feats[name] = x
    c3, c4, c5 = feats["blocks4"], feats["blocks5"], feats["blocks6"]
    # ---- CSP-PAN (csp_pan.py:310-347).  Concatenation buffers: cat_a = [up(t2) | t1], cat_b = [up(inner1) | t0],
    # cat_c = [down0(inner0) | inner1], cat_d = [down1(out1) | t2]
    c = PICO_NECK_CH
    cat_a, cat_b, cat_c, cat_d = b.tensor(2 * c, 16), b.tensor(2 * c, 8), b.tensor(2 * c, 16), b.tensor(2 * c, 32)

    def conv_t(i, src, dst, coff):
        b.op(OP_PW, src, dst, out_coff=coff, out_c=c, act=ACT_HSWISH, w=_pw(b, neck, f"conv_t.convs.{i}.conv.weight", f"conv_t.convs.{i}.bn"))

    conv_t(0, c3, cat_b, c)
    conv_t(1, c4, cat_a, c)
    conv_t(2, c5, cat_d, c)

    def dp(p, src, src_coff, ch, dst, dst_coff, stride, down_out):
        t = b.tensor(ch, down_out)
        wid, kk = _dw(b, neck, p + ".dwconv.weight", p + ".bn1")
        b.op(OP_DW, src, t, in_coff=src_coff, in_c=ch, k=kk, stride=stride, act=ACT_HSWISH, w=wid)
        b.op(OP_PW, t, dst, out_coff=dst_coff, out_c=ch, act=ACT_HSWISH, w=_pw(b, neck, p + ".pwconv.weight", p + ".bn2"))

    def csp(p, cat, down_l, dst, dst_coff):
        mid = c // 2
        cat2 = b.tensor(2 * mid, down_l)  # [main | short]
        b.op(OP_PW, cat, cat2, out_coff=mid, out_c=mid, act=ACT_HSWISH, w=_pw(b, neck, p + ".short_conv.conv.weight", p + ".short_conv.bn"))
        m0 = b.tensor(mid, down_l)
        b.op(OP_PW, cat, m0, act=ACT_HSWISH, w=_pw(b, neck, p + ".main_conv.conv.weight", p + ".main_conv.bn"))
        m1 = b.tensor(mid, down_l)
        b.op(OP_PW, m0, m1, act=ACT_HSWISH, w=_pw(b, neck, p + ".blocks.0.conv1.conv.weight", p + ".blocks.0.conv1.bn"))
        dp(p + ".blocks.0.conv2", m1, 0, mid, cat2, 0, 1, down_l)
        b.op(OP_PW, cat2, dst, out_coff=dst_coff, out_c=c, act=ACT_HSWISH, w=_pw(b, neck, p + ".final_conv.conv.weight", p + ".final_conv.bn"))

    b.op(OP_UP2, cat_d, cat_a, in_coff=c, in_c=c, out_coff=0, out_c=c)          # up(t2) -> cat_a[:c]
    csp("top_down_blocks.0", cat_a, 16, cat_c, c)                                # inner1 -> cat_c[c:]
    b.op(OP_UP2, cat_c, cat_b, in_coff=c, in_c=c, out_coff=0, out_c=c)          # up(inner1) -> cat_b[:c]
    p3 = b.tensor(c, 8)
    csp("top_down_blocks.1", cat_b, 8, p3, 0)                                    # inner0 = level-0 output
    dp("downsamples.0", p3, 0, c, cat_c, 0, 2, 16)
    p4 = b.tensor(c, 16)
    csp("bottom_up_blocks.0", cat_c, 16, p4, 0)
    dp("downsamples.1", p4, 0, c, cat_d, 0, 2, 32)
    p5 = b.tensor(c, 32)
    csp("bottom_up_blocks.1", cat_d, 32, p5, 0)
    top_a, top_b, p6 = b.tensor(c, 64), b.tensor(c, 64), b.tensor(c, 64)
    dp("first_top_conv", cat_d, c, c, top_a, 0, 2, 64)                           # inputs[-1] = t2 lives in cat_d[c:]
    dp("second_top_conv", p5, 0, c, top_b, 0, 2, 64)
    b.op(OP_ADD, top_a, p6, aux=top_b)
    # ---- PicoHead (pico_head.py:151-167, 1108-1138)
    n_out = num_classes + 32
    pad_out = (n_out + 7) // 8 * 8
    for lvl, (x, d) in enumerate(((p3, 8), (p4, 16), (p5, 32), (p6, 64))):
        for i in range(PICO_HEAD_CONVS):
            t = b.tensor(c, d)
            wid, kk = _dw(b, head, f"conv_feat.cls_conv_dw{lvl}_{i}.conv.weight", f"conv_feat.cls_conv_dw{lvl}_{i}.norm")
            b.op(OP_DW, x, t, k=kk, stride=1, act=ACT_HSWISH, w=wid)
            x = b.tensor(c, d)
            b.op(OP_PW, t, x, act=ACT_HSWISH, w=_pw(b, head, f"conv_feat.cls_conv_pw{lvl}_{i}.conv.weight", f"conv_feat.cls_conv_pw{lvl}_{i}.norm"))
        hw = np.zeros((pad_out, c, 1, 1), np.float32)
        hb = np.zeros(pad_out, np.float32)
        hw[:n_out] = _f(head, f"head_cls{lvl}.weight")
        hb[:n_out] = _f(head, f"head_cls{lvl}.bias")
        wp, bp = W.pack_conv(hw, hb)
        b.op(OP_HEAD, x, x, out_c=pad_out, act=ACT_NONE, w=b.weight(w=wp, b=bp), aux=lvl)
    blob = dict(b.blob)
    blob["graph.tensors"] = np.array(b.tensors, np.int32)
    blob["graph.ops"] = np.array(b.ops, np.int32)
    blob["graph.meta"] = np.array([num_classes, 32, pad_out, len(b.tensors), len(b.ops), 0, 0, 0], np.int32)
    return blob, {"features": {"c3": c3, "c4": c4, "c5": c5, "p3": p3, "p4": p4, "p5": p5, "p6": p6}}


def pack_picodet(backbone: Mapping, neck: Mapping, head: Mapping, num_classes: int) -> bytes:
    blob, _ = build_picodet(backbone, neck, head, num_classes)
    return W.write_blob(blob)
