"""PP-OCRv4 text recogniser ("SVTR-LCNet": PPLCNetV3-0.95 backbone -> SVTR neck -> CTC head) as a graph program
for the engine's executor (csrc/graph_net.cu), model kind "pp_rec".

The reference runs this network as an ONNX file from the hub (ocr_pdf/ocr_table_model_config.py:166-204, executed at
ocr_pdf/ocr_recognition_task.py:90-99); the architecture lowered here is the published one that file was exported from
(PaddleOCR release 2.7: rec_lcnetv3.py deploy form, necks/rnn.py EncoderWithSVTR, heads/rec_ctc_head.py -- restated in
oracle/pp_rec_ref.py, whose state-dict keys this module consumes: the Paddle module tree with torch conventions).

Lowering rules (checked on CPU by running the program with oracle/graph_interp.py against the oracle, and on the GPU against
the oracle's probabilities):
  * a rep layer = conv + bias -> lab (scalar scale s1, bias c1) -> hardswish -> lab (s2, c2): s1 / c1 are folded into the
    conv ((s1 W) x + (s1 b + c1)); s2 / c2 stay an explicit post-activation affine of the op (`w{id}.pa`) because the next
    layer is a zero-padded depthwise conv or an SE gate, across which a bias cannot be folded;
  * conv1 (3x3 s2, BatchNorm, no activation) is the executor's stem kernel with the PP rec normalisation fused;
  * the (1,3) convs of the neck are OP_UNFOLD3 + a flat GEMM (K = 3 Cin, k = tap * Cin + c); 60-channel tensors are padded to 64;
  * BatchNorm folded into the neck convs; the attention scale head_dim ** -0.5 folded into the q rows of the qkv projection;
  * the CTC head's Linear is padded to a multiple of 8 classes (zero rows; the softmax kernel reads the first n_class columns).
"""
from __future__ import annotations

from typing import Dict, List, Mapping, Tuple

import numpy as np

from . import weights as W
from .picodet_graph import OP_DW, OP_PW, OP_SE, OP_STEM
from .synth import PP_REC_CONFIG, pp_rec_ch

OP_AVGPOOL, OP_UNFOLD3, OP_LN, OP_ATTN, OP_CTC = 7, 8, 9, 10, 11
ACT_NONE, ACT_HSWISH, ACT_SWISH = 0, 4, 5
NECK_HEADS = 8


class _Builder:
    def __init__(self):
        self.tensors: List[Tuple[int, int, int, int, int]] = []
        self.ops: List[List[int]] = []
        self.blob: Dict[str, np.ndarray] = {}
        self.nw = 0

    def tensor(self, c: int, dh: int, dw: int, ph: int = 1, pw: int = 1) -> int:
        self.tensors.append((c, dh, dw, ph, pw))
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


def _f(sd, k) -> np.ndarray:
    return W._np(sd[k]).astype(np.float32)


def _pad_to(a: np.ndarray, axis: int, n: int) -> np.ndarray:
    if a.shape[axis] == n:
        return a
    pad = [(0, 0)] * a.ndim
    pad[axis] = (0, n - a.shape[axis])
    return np.pad(a, pad)


def pw_pack_factor(b: _Builder, cin: int, cout: int) -> int:
    """Pixels per GEMM row for a narrow 1x1 layer (the executor's OP_PW with k = factor): a 16- or 32-channel layer is run on
    4 (2) consecutive pixels at once against a block-diagonal weight -- K = 64 (128-byte TMA rows instead of 32 / 64-byte ones),
    N = factor * cout -- because at K = 16 a 128-pixel tile is one MMA and the kernel is bound by per-tile TMA rows and barrier
    round trips, not by bytes.  fp32x programs keep one pixel per row."""
    if getattr(b, "precise", False) or cin > 32 or cin % 16 or cout * (64 // cin) > 256:
        return 1
    return 64 // cin


def block_diagonal(w2d: np.ndarray, bias: np.ndarray, pack: int):
    """w [Cout, K] -> diag(w, ..., w) [pack * Cout, pack * K] and the bias tiled: output row j * Cout + o of a packed row is
    channel o of its j-th pixel."""
    cout, k = w2d.shape
    out = np.zeros((pack * cout, pack * k), np.float32)
    for j in range(pack):
        out[j * cout:(j + 1) * cout, j * k:(j + 1) * k] = w2d
    return out, np.tile(bias, pack)


def _gemm_weight(b: _Builder, w2d: np.ndarray, bias: np.ndarray, pa=None, pack: int = 1) -> int:
    """[Cout, K] fp32 (+bias) -> packed 1x1 weight of the executor (fp32x builders: the split-fp16 triple [W_hi | W_lo | W_hi])."""
    if pack > 1:
        w2d, bias = block_diagonal(w2d, bias, pack)
    if not getattr(b, "precise", False) and w2d.shape[1] >= 128 and w2d.shape[1] % 64:
        # wide layers: K padded to whole 64-channel k-blocks (240 -> 256, 480 -> 512); the executor's TMA zero-fills the A columns
        w2d = _pad_to(w2d, 1, (w2d.shape[1] + 63) // 64 * 64)
    wp, bp = W.pack_split_linear(w2d, bias) if getattr(b, "precise", False) else W.pack_conv(w2d[:, :, None, None], bias)
    arrays = {"w": wp, "b": bp}
    if pa is not None:
        arrays["pa"] = np.array(pa, np.float32)
    return b.weight(**arrays)


def build_pp_rec(sd: Mapping, n_class: int = None, precise: bool = False):
    """-> (blob tensor dict for weights.write_blob, meta).  Program format: picodet_graph.py's, with the 5-column tensor table
    (channels, down_h, down_w, pool_h, pool_w) and the extra opcodes of csrc/graph_net.cu.  precise=True: the fp32x mode -- every
    GEMM weight as a split-fp16 triple and a "precision" entry that makes the executor keep its activations in fp32."""
    b = _Builder()
    b.precise = precise
    n_class = int(_f(sd, "head.ctc_head.fc.weight").shape[0]) if n_class is None else n_class
    img = b.tensor(3, 1, 1)
    # ---- conv1: 3x3 stride 2 + BatchNorm, no activation
    w = _f(sd, "backbone.conv1.conv.weight")
    scale, shift = W.bn_affine({k: _f(sd, f"backbone.conv1.bn.{k}") for k in ("weight", "bias", "running_mean", "running_var")}, 16)
    dh = dw = 2
    x = b.tensor(16, dh, dw)
    b.op(OP_STEM, img, x, k=3, stride=2, act=ACT_NONE,
         w=b.weight(sw=(w * scale[:, None, None, None]).transpose(2, 3, 1, 0).reshape(27, 16).astype(np.float32), sb=shift.astype(np.float32)))

    def lab(p):
        return float(_f(sd, p + ".scale")[0]), float(_f(sd, p + ".bias")[0])

    feats = {}
    for name, cfg in PP_REC_CONFIG.items():
        for i, (k, cin, cout, s, se) in enumerate(cfg):
            ci, co = pp_rec_ch(cin), pp_rec_ch(cout)
            p = f"backbone.{name}.{i}"
            sh, sw = (s, s) if isinstance(s, int) else s
            dh, dw = dh * sh, dw * sw
            # depthwise rep layer
            s1, c1 = lab(p + ".dw_conv.lab")
            wd = _f(sd, p + ".dw_conv.reparam_conv.weight")[:, 0] * s1  # [C,k,k]
            bd = _f(sd, p + ".dw_conv.reparam_conv.bias") * s1 + c1
            t = b.tensor(ci, dh, dw)
            b.op(OP_DW, x, t, k=k, stride=sh if sh == sw else (sh | (sw << 8)), act=ACT_HSWISH,
                 w=b.weight(dw=wd.transpose(1, 2, 0).reshape(k * k, ci).astype(np.float32), db=bd.astype(np.float32),
                            pa=np.array(lab(p + ".dw_conv.act.lab"), np.float32)))
            if se:
                t2 = b.tensor(ci, dh, dw)
                b.op(OP_SE, t, t2, w=b.weight(s1w=_f(sd, p + ".se.conv1.weight").reshape(ci // 4, ci), s1b=_f(sd, p + ".se.conv1.bias"),
                                             s2w=_f(sd, p + ".se.conv2.weight").reshape(ci, ci // 4), s2b=_f(sd, p + ".se.conv2.bias")))
                t = t2
            # pointwise rep layer
            s1, c1 = lab(p + ".pw_conv.lab")
            wp = _f(sd, p + ".pw_conv.reparam_conv.weight").reshape(co, ci) * s1
            bp = _f(sd, p + ".pw_conv.reparam_conv.bias") * s1 + c1
            x = b.tensor(co, dh, dw)
            pk = pw_pack_factor(b, ci, co)
            b.op(OP_PW, t, x, k=pk, act=ACT_HSWISH, w=_gemm_weight(b, wp, bp, pa=lab(p + ".pw_conv.act.lab"), pack=pk))
        feats[name] = x
    cb = pp_rec_ch(512)
    d = int(_f(sd, "head.ctc_encoder.encoder.norm.weight").shape[0])
    c8 = cb // 8
    c8p = (c8 + 7) // 8 * 8
    # ---- eval tail avg_pool2d(x, [3, 2]) written straight into the neck's concatenation buffer [h | conv3(z)]
    line = (dh, dw, 3, 2)
    cat = b.tensor(2 * cb, *line)
    b.op(OP_AVGPOOL, x, cat, out_coff=0, out_c=cb, k=3 | (2 << 8))
    q = "head.ctc_encoder.encoder"

    def bn_fold(name, cout):
        return W.bn_affine({k: _f(sd, f"{q}.{name}.norm.{k}") for k in ("weight", "bias", "running_mean", "running_var")}, cout)

    def conv13(name, src, src_coff, cin, cout, cout_pad):
        """(1,3) conv + BN + Swish as unfold + GEMM; K index = tap * cin + c."""
        w = _f(sd, f"{q}.{name}.conv.weight")  # [cout, cin, 1, 3]
        scale, shift = bn_fold(name, cout)
        w2 = (w[:, :, 0, :] * scale[:, None, None]).transpose(0, 2, 1).reshape(cout, 3 * cin)
        u = b.tensor(3 * cin, *line)
        b.op(OP_UNFOLD3, src, u, in_coff=src_coff, in_c=cin)
        out = b.tensor(cout_pad, *line)
        b.op(OP_PW, u, out, act=ACT_SWISH, w=_gemm_weight(b, _pad_to(w2, 0, cout_pad), _pad_to(shift, 0, cout_pad)))
        return out

    def conv11(name, src, cin, cin_pad, cout, dst, dst_coff=0, res=-1):
        w = _f(sd, f"{q}.{name}.conv.weight").reshape(cout, cin)
        scale, shift = bn_fold(name, cout)
        b.op(OP_PW, src, dst, out_coff=dst_coff, out_c=cout, act=ACT_SWISH, w=_gemm_weight(b, _pad_to(w * scale[:, None], 1, cin_pad), shift), aux=res)

    z = conv13("conv1", cat, 0, cb, c8, c8p)
    zs = b.tensor(d, *line)
    conv11("conv2", z, c8, c8p, d, zs)

    def linear(name, src, cin, cout, dst, act=ACT_NONE, res=-1, row_scale=None):
        w, bias = _f(sd, name + ".weight"), _f(sd, name + ".bias")
        if row_scale is not None:
            w, bias = w * row_scale[:, None], bias * row_scale
        b.op(OP_PW, src, dst, act=act, w=_gemm_weight(b, w, bias), aux=res)

    def ln(name, src, eps):
        out = b.tensor(d, *line)
        b.op(OP_LN, src, out, w=b.weight(lnw=_f(sd, name + ".weight"), lnb=_f(sd, name + ".bias"), eps=np.array([eps], np.float32)))
        return out

    hd = d // NECK_HEADS
    qscale = np.concatenate([np.full(d, hd ** -0.5, np.float32), np.ones(2 * d, np.float32)])
    i = 0
    while f"{q}.svtr_block.{i}.norm1.weight" in sd:
        p = f"{q}.svtr_block.{i}"
        h = ln(p + ".norm1", zs, 1e-5)
        qkv = b.tensor(3 * d, *line)
        linear(p + ".mixer.qkv", h, d, 3 * d, qkv, row_scale=qscale)
        ctx = b.tensor(d, *line)
        b.op(OP_ATTN, qkv, ctx, k=NECK_HEADS)
        z1 = b.tensor(d, *line)
        linear(p + ".mixer.proj", ctx, d, d, z1, res=zs)
        h = ln(p + ".norm2", z1, 1e-5)
        m1 = b.tensor(int(_f(sd, p + ".mlp.fc1.weight").shape[0]), *line)
        linear(p + ".mlp.fc1", h, d, b.tensors[m1][0], m1, act=ACT_SWISH)
        zs = b.tensor(d, *line)
        linear(p + ".mlp.fc2", m1, b.tensors[m1][0], d, zs, res=z1)
        i += 1
    zn = ln(q + ".norm", zs, 1e-6)
    conv11("conv3", zn, d, d, cb, cat, dst_coff=cb)
    z = conv13("conv4", cat, 0, 2 * cb, c8, c8p)
    seq = b.tensor(d, *line)
    conv11("conv1x1", z, c8, c8p, d, seq)
    # ---- CTC head
    pad_out = (n_class + 7) // 8 * 8
    fw, fb = _f(sd, "head.ctc_head.fc.weight"), _f(sd, "head.ctc_head.fc.bias")
    b.op(OP_CTC, seq, seq, out_c=pad_out, w=_gemm_weight(b, _pad_to(fw, 0, pad_out), _pad_to(fb, 0, pad_out)))
    blob = dict(b.blob)
    if precise:
        blob["precision"] = np.array([1], np.int32)
    blob["graph.tensors"] = np.array(b.tensors, np.int32)
    blob["graph.ops"] = np.array(b.ops, np.int32)
    blob["graph.meta"] = np.array([n_class, 0, pad_out, len(b.tensors), len(b.ops), 1, 0, 0], np.int32)
    return blob, {"features": feats, "pooled": cat, "seq": seq, "n_class": n_class}


def pack_pp_rec(sd: Mapping, n_class: int = None, precise: bool = False) -> bytes:
    blob, _ = build_pp_rec(sd, n_class, precise)
    return W.write_blob(blob)
