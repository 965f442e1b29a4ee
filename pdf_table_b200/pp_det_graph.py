"""PP-OCRv4 mobile text detector (PPLCNetV3-0.75 backbone -> RSE-FPN -> DBHead) as a graph program for the
engine's executor (csrc/graph_net.cu), model kind "pp_det".

The reference runs this network as an ONNX file from the hub (ocr_pdf/ocr_table_model_config.py:134-147, executed at
ocr_pdf/ocr_detection_task.py:98-107); the architecture lowered here is the published one that file was exported from
(PaddleOCR release 2.7: rec_lcnetv3.py det=True deploy form, necks/db_fpn.py RSEFPN, heads/det_db_head.py -- restated in
oracle/pp_det_ref.py, whose state-dict keys this module consumes).

Lowering rules (checked on CPU by running the program with oracle/graph_interp.py against the oracle, and on the GPU against
the oracle's probability map):
  * rep layers as in pp_rec_graph.py (first affine folded into the conv, second kept as the op's post-activation affine); a
    stride-2 rep layer has NO activation and no second affine (LearnableRepLayer.forward skips `act` for stride 2);
  * each backbone tap conv (1x1 with bias, to 12 / 18 / 42 / 360 channels) and the neck's ins_conv (1x1, no bias, to 96) are
    two linear maps with nothing between them: folded into ONE 1x1 conv W = W_ins W_tap, b = W_ins b_tap, so the odd-width tap
    tensors never exist;
  * RSELayer = conv -> x + x * s: the executor's SE op with the shortcut flag (k = 2), i.e. x * (1 + s); the neck's
    hardsigmoid has slope 0.2 (the backbone's 1/6): hardsigmoid_0.2(z) = hardsigmoid_1/6(1.2 z), folded into conv2;
  * the top-down path out_k = in_k + up2(out_{k+1}) is OP_UP2 with an addend; p5 / p4 / p3 are up-sampled (nearest x8 / x4 /
    x2) straight into their slices of the 96-channel concatenation, p2's SE writes its slice directly;
  * head: conv3x3 + BN + relu as a dense-conv op (output padded 24 -> 32 channels with zero filters so that the transposed
    conv's K is one 32-wide block); ConvTranspose 2x2 s2 + BN + relu as a GEMM with the pixel-shuffle store (4 x 32 padded
    output columns); the last ConvTranspose 24 -> 1 + sigmoid writes the fp32 probability map.
"""
from __future__ import annotations

from typing import Mapping

import numpy as np

from . import weights as W
from .picodet_graph import OP_DW, OP_PW, OP_SE, OP_STEM, OP_UP2
from .pp_rec_graph import ACT_HSWISH, ACT_NONE, _Builder, _f, _gemm_weight, _pad_to, pw_pack_factor
from .synth import PP_DET_CONFIG, PP_DET_FPN, PP_DET_SCALE, pp_rec_ch

OP_CONV, OP_DECONV2, OP_DBHEAD = 12, 13, 14
ACT_RELU = 1
HEAD_PAD = 32  # the 24-channel head tensors are padded to 32


def build_pp_det(sd: Mapping, precise: bool = False):
    """-> (blob tensor dict for weights.write_blob, meta).  Program format: pp_rec_graph.py's.  precise=True: the fp32x mode (fp32
    activation buffers, every GEMM / conv weight as a split-fp16 triple, see csrc/graph_net.cu GraphNet::precise)."""
    b = _Builder()
    b.precise = precise
    pack3x3 = W.pack_conv_split if precise else W.pack_conv
    ch = lambda c: pp_rec_ch(c, PP_DET_SCALE)  # noqa: E731
    img = b.tensor(3, 1, 1)
    w = _f(sd, "backbone.conv1.conv.weight")
    scale, shift = W.bn_affine({k: _f(sd, f"backbone.conv1.bn.{k}") for k in ("weight", "bias", "running_mean", "running_var")}, 16)
    d = 2
    x = b.tensor(16, d, d)
    b.op(OP_STEM, img, x, k=3, stride=2, act=ACT_NONE,
         w=b.weight(sw=(w * scale[:, None, None, None]).transpose(2, 3, 1, 0).reshape(27, 16).astype(np.float32), sb=shift.astype(np.float32)))

    def lab(p):
        return float(_f(sd, p + ".scale")[0]), float(_f(sd, p + ".bias")[0])

    taps = []
    for name, cfg in PP_DET_CONFIG.items():
        for i, (k, cin, cout, s, se) in enumerate(cfg):
            ci, co = ch(cin), ch(cout)
            p = f"backbone.{name}.{i}"
            d *= s
            s1, c1 = lab(p + ".dw_conv.lab")
            wd = _f(sd, p + ".dw_conv.reparam_conv.weight")[:, 0] * s1
            bd = _f(sd, p + ".dw_conv.reparam_conv.bias") * s1 + c1
            arrays = dict(dw=wd.transpose(1, 2, 0).reshape(k * k, ci).astype(np.float32), db=bd.astype(np.float32))
            if s != 2:
                arrays["pa"] = np.array(lab(p + ".dw_conv.act.lab"), np.float32)
            t = b.tensor(ci, d, d)
            b.op(OP_DW, x, t, k=k, stride=s, act=ACT_HSWISH if s != 2 else ACT_NONE, w=b.weight(**arrays))
            if se:
                t2 = b.tensor(ci, d, d)
                b.op(OP_SE, t, t2, w=b.weight(s1w=_f(sd, p + ".se.conv1.weight").reshape(ci // 4, ci), s1b=_f(sd, p + ".se.conv1.bias"),
                                             s2w=_f(sd, p + ".se.conv2.weight").reshape(ci, ci // 4), s2b=_f(sd, p + ".se.conv2.bias")))
                t = t2
            s1, c1 = lab(p + ".pw_conv.lab")
            wp = _f(sd, p + ".pw_conv.reparam_conv.weight").reshape(co, ci) * s1
            bp = _f(sd, p + ".pw_conv.reparam_conv.bias") * s1 + c1
            x = b.tensor(co, d, d)
            pk = pw_pack_factor(b, ci, co)
            b.op(OP_PW, t, x, k=pk, act=ACT_HSWISH, w=_gemm_weight(b, wp, bp, pa=lab(p + ".pw_conv.act.lab"), pack=pk))
        if name != "blocks2":
            taps.append((x, co, d))

    f, q = PP_DET_FPN, PP_DET_FPN // 4

    def se_weights(p, c):
        # neck hardsigmoid: clip(0.2 z + 0.5, 0, 1) = clip((1.2 z) / 6 + 0.5, 0, 1)
        return b.weight(s1w=_f(sd, p + ".conv1.weight").reshape(c // 4, c), s1b=_f(sd, p + ".conv1.bias"),
                        s2w=_f(sd, p + ".conv2.weight").reshape(c, c // 4) * np.float32(1.2), s2b=_f(sd, p + ".conv2.bias") * np.float32(1.2))

    # ---- ins_conv (tap conv folded in) + RSE
    ins = []
    for i, (t, c, dd) in enumerate(taps):
        wt, bt = _f(sd, f"backbone.layer_list.{i}.weight"), _f(sd, f"backbone.layer_list.{i}.bias")
        wi = _f(sd, f"neck.ins_conv.{i}.in_conv.weight")
        w2 = wi.reshape(f, -1).astype(np.float64) @ wt.reshape(wt.shape[0], c).astype(np.float64)
        b2 = wi.reshape(f, -1).astype(np.float64) @ bt.astype(np.float64)
        u = b.tensor(f, dd, dd)
        b.op(OP_PW, t, u, act=ACT_NONE, w=_gemm_weight(b, w2.astype(np.float32), b2.astype(np.float32)))
        o = b.tensor(f, dd, dd)
        b.op(OP_SE, u, o, k=2, w=se_weights(f"neck.ins_conv.{i}.se_block", f))
        ins.append((o, dd))
    # ---- top-down sums
    outs = [None, None, None, ins[3]]
    for i in (2, 1, 0):
        o = b.tensor(f, ins[i][1], ins[i][1])
        b.op(OP_UP2, outs[i + 1][0], o, aux=ins[i][0])
        outs[i] = (o, ins[i][1])
    # ---- inp_conv + RSE, up-sampled into the concatenation [p5 | p4 | p3 | p2]
    d4 = taps[0][2]
    fuse = b.tensor(f, d4, d4)
    for i in range(4):
        src, dd = outs[i]
        wc = _f(sd, f"neck.inp_conv.{i}.in_conv.weight")  # [24, 96, 3, 3]
        u = b.tensor(q, dd, dd)
        wp, bp = pack3x3(wc, np.zeros(q, np.float32))
        b.op(OP_CONV, src, u, k=3, stride=1, act=ACT_NONE, w=b.weight(w=wp, b=bp))
        coff = (3 - i) * q
        if i == 0:
            b.op(OP_SE, u, fuse, out_coff=coff, out_c=q, k=2, w=se_weights(f"neck.inp_conv.{i}.se_block", q))
        else:
            o = b.tensor(q, dd, dd)
            b.op(OP_SE, u, o, k=2, w=se_weights(f"neck.inp_conv.{i}.se_block", q))
            b.op(OP_UP2, o, fuse, out_coff=coff, out_c=q)
    # ---- DBHead.binarize
    p = "head.binarize"
    bn1 = {k: _f(sd, f"{p}.conv_bn1.{k}") for k in ("weight", "bias", "running_mean", "running_var")}
    wp, bp = pack3x3(_pad_to(_f(sd, p + ".conv1.weight"), 0, HEAD_PAD), None,
                         {k: np.concatenate([v, np.ones(HEAD_PAD - q, np.float32) if k in ("weight", "running_var") else np.zeros(HEAD_PAD - q, np.float32)])
                          for k, v in bn1.items()})
    bp[q:HEAD_PAD] = 0.0
    h1 = b.tensor(HEAD_PAD, d4, d4)
    b.op(OP_CONV, fuse, h1, k=3, stride=1, act=ACT_RELU, w=b.weight(w=wp, b=bp))
    bn2 = {k: _f(sd, f"{p}.conv_bn2.{k}") for k in ("weight", "bias", "running_mean", "running_var")}
    w2 = np.zeros((HEAD_PAD, HEAD_PAD, 2, 2), np.float32)  # ConvTranspose weight [Cin, Cout, 2, 2]
    w2[:q, :q] = _f(sd, p + ".conv2.weight")
    b2 = np.zeros(HEAD_PAD, np.float32)
    b2[:q] = _f(sd, p + ".conv2.bias")
    bn2p = {k: np.concatenate([v, np.ones(HEAD_PAD - q, np.float32) if k in ("weight", "running_var") else np.zeros(HEAD_PAD - q, np.float32)])
            for k, v in bn2.items()}
    wp, bp = W.pack_deconv2x2(w2, b2, bn2p, split=precise)
    h2 = b.tensor(HEAD_PAD, d4 // 2, d4 // 2)
    b.op(OP_DECONV2, h1, h2, act=ACT_RELU, w=b.weight(w=wp, b=bp))
    w3 = np.zeros((HEAD_PAD, 4), np.float32)
    w3[:q] = _f(sd, p + ".conv3.weight").reshape(q, 4)  # [c][dy * 2 + dx]
    b.op(OP_DBHEAD, h2, h2, w=b.weight(hw=w3, hb=_f(sd, p + ".conv3.bias").reshape(1)))
    blob = dict(b.blob)
    if precise:
        blob["precision"] = np.array([1], np.int32)
    blob["graph.tensors"] = np.array(b.tensors, np.int32)
    blob["graph.ops"] = np.array(b.ops, np.int32)
    blob["graph.meta"] = np.array([1, 0, 8, len(b.tensors), len(b.ops), 3, 0, 0], np.int32)
    return blob, {"fuse": fuse, "ins": [t for t, _ in ins], "outs": [t for t, _ in outs], "h1": h1, "h2": h2}


def pack_pp_det(sd: Mapping, precise: bool = False) -> bytes:
    blob, _ = build_pp_det(sd, precise)
    return W.write_blob(blob)
