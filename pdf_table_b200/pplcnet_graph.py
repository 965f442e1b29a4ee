"""PULC image classifiers (PP-LCNet x1.0) as a graph program for the engine's executor (csrc/graph_net.cu),
model kind "pplcnet_cls".

Lowers a state_dict of the reference module ``PPLCNet`` (cls/cls_pp_lcnet.py:164-293: conv1 -> blocks2..6 of DepthwiseSeparable
(dw ConvBNLayer -> [SEModule] -> pw ConvBNLayer, hardswish after every BatchNorm) -> global average pool -> last_conv 1x1
(512 -> 1280, no bias) + hardswish -> fc) for one of the reference's tasks (cls/configuration_cls_pulc.py:20-42): the stride list
(all 2, or (2,1) for the text-line tasks) is baked into the program.  BatchNorm is folded; the op set is PicoDet's LCNet one plus
the global pool and the softmax head of the recogniser.
"""
from __future__ import annotations

from typing import Mapping, Sequence

import numpy as np

from . import weights as W
from .picodet_graph import OP_DW, OP_PW, OP_SE, OP_STEM
from .pp_rec_graph import ACT_HSWISH, ACT_NONE, OP_AVGPOOL, OP_CTC, _Builder, _f, _gemm_weight, _pad_to, pw_pack_factor

NET_CONFIG = {  # cls/cls_pp_lcnet.py:52-63: k, in_c, out_c, stride, use_se (the first stride of blocks3..6 comes from stride_list)
    "blocks2": [[3, 16, 32, 1, False]],
    "blocks3": [[3, 32, 64, 2, False], [3, 64, 64, 1, False]],
    "blocks4": [[3, 64, 128, 2, False], [3, 128, 128, 1, False]],
    "blocks5": [[3, 128, 256, 2, False], [5, 256, 256, 1, False], [5, 256, 256, 1, False], [5, 256, 256, 1, False], [5, 256, 256, 1, False],
                [5, 256, 256, 1, False]],
    "blocks6": [[5, 256, 512, 2, True], [5, 512, 512, 1, True]],
}
TASK_STRIDES = {  # cls/configuration_cls_pulc.py CLS_PULC_TASK_CONFIG
    "table_attribute": [2, 2, 2, 2, 2], "text_image_orientation": [2, 2, 2, 2, 2],
    "textline_orientation": [2, (2, 1), (2, 1), (2, 1), (2, 1)], "language_classification": [2, (2, 1), (2, 1), (2, 1), (2, 1)],
}
TASK_CLASSES = {"table_attribute": 6, "text_image_orientation": 4, "textline_orientation": 2, "language_classification": 10}


def net_config(stride_list: Sequence) -> dict:
    cfg = {k: [list(r) for r in v] for k, v in NET_CONFIG.items()}
    for i, s in enumerate(stride_list[1:]):
        cfg[f"blocks{i + 3}"][0][3] = s
    return cfg


def build_pplcnet(sd: Mapping, stride_list: Sequence = (2, 2, 2, 2, 2)):
    b = _Builder()
    class_num = int(_f(sd, "fc.weight").shape[0])

    def bn(p, c):
        return W.bn_affine({k: _f(sd, f"{p}.{k}") for k in ("weight", "bias", "running_mean", "running_var")}, c)

    img = b.tensor(3, 1, 1)
    w = _f(sd, "conv1.conv.weight")
    scale, shift = bn("conv1.bn", 16)
    dh = dw = 2
    x = b.tensor(16, dh, dw)
    b.op(OP_STEM, img, x, k=3, stride=2, act=ACT_HSWISH,
         w=b.weight(sw=(w * scale[:, None, None, None]).transpose(2, 3, 1, 0).reshape(27, 16).astype(np.float32), sb=shift.astype(np.float32)))
    for name, cfg in net_config(stride_list).items():
        for i, (k, ci, co, s, se) in enumerate(cfg):
            p = f"{name}.{i}"
            sh, sw = (s, s) if isinstance(s, int) else s
            dh, dw = dh * sh, dw * sw
            wd = _f(sd, p + ".dw_conv.conv.weight")[:, 0]
            scale, shift = bn(p + ".dw_conv.bn", ci)
            t = b.tensor(ci, dh, dw)
            b.op(OP_DW, x, t, k=k, stride=sh if sh == sw else (sh | (sw << 8)), act=ACT_HSWISH,
                 w=b.weight(dw=(wd * scale[:, None, None]).transpose(1, 2, 0).reshape(k * k, ci).astype(np.float32), db=shift.astype(np.float32)))
            if se:
                t2 = b.tensor(ci, dh, dw)
                b.op(OP_SE, t, t2, w=b.weight(s1w=_f(sd, p + ".se.conv1.weight").reshape(ci // 4, ci), s1b=_f(sd, p + ".se.conv1.bias"),
                                             s2w=_f(sd, p + ".se.conv2.weight").reshape(ci, ci // 4), s2b=_f(sd, p + ".se.conv2.bias")))
                t = t2
            scale, shift = bn(p + ".pw_conv.bn", co)
            x = b.tensor(co, dh, dw)
            pk = pw_pack_factor(b, ci, co)
            b.op(OP_PW, t, x, k=pk, act=ACT_HSWISH, w=_gemm_weight(b, _f(sd, p + ".pw_conv.conv.weight").reshape(co, ci) * scale[:, None], shift, pack=pk))
    pooled = b.tensor(512, dh, dw, 0, 0)
    b.op(OP_AVGPOOL, x, pooled, k=0)
    expand = int(_f(sd, "last_conv.weight").shape[0])
    feat = b.tensor(expand, dh, dw, 0, 0)
    b.op(OP_PW, pooled, feat, act=ACT_HSWISH, w=_gemm_weight(b, _f(sd, "last_conv.weight").reshape(expand, 512), np.zeros(expand, np.float32)))
    pad_out = (class_num + 7) // 8 * 8
    b.op(OP_CTC, feat, feat, out_c=pad_out, act=ACT_NONE, w=_gemm_weight(b, _pad_to(_f(sd, "fc.weight"), 0, pad_out), _pad_to(_f(sd, "fc.bias"), 0, pad_out)))
    blob = dict(b.blob)
    blob["graph.tensors"] = np.array(b.tensors, np.int32)
    blob["graph.ops"] = np.array(b.ops, np.int32)
    blob["graph.meta"] = np.array([class_num, 0, pad_out, len(b.tensors), len(b.ops), 2, 0, 0], np.int32)
    return blob, {"class_num": class_num}


def pack_pplcnet(sd: Mapping, stride_list: Sequence = (2, 2, 2, 2, 2)) -> bytes:
    return W.write_blob(build_pplcnet(sd, stride_list)[0])
