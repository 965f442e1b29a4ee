"""Weight packer: reference state_dict -> flat blob consumed by libdocvision.so (dv_create).

This is the engine's only "checkpoint" format: BatchNorm folded into the
convolution, weights converted to fp16 and laid out K-major as [Cout][tap][Cin_pad] so one TMA box is
one (tap, channel-block) K-slice of the implicit GEMM; biases stay fp32 and are padded to 256.

Blob layout (little endian; mirrored in csrc/capi.cu):
    header  : magic "DVWBLOB1", u32 n_tensors, u32 reserved, u64 data_offset, u64 data_bytes
    entries : n_tensors x { char name[96]; u32 dtype; u32 ndim; u32 dims[4]; u64 offset; u64 nbytes }
    payload : tensors, each 256-byte aligned
"""
from __future__ import annotations

import struct
from typing import Dict, Mapping, Optional, Tuple

import numpy as np

DT_F32, DT_F16, DT_I32 = 0, 1, 2
_DT = {np.dtype(np.float32): DT_F32, np.dtype(np.float16): DT_F16, np.dtype(np.int32): DT_I32}


def _np(t) -> np.ndarray:
    if hasattr(t, "detach"):
        t = t.detach().cpu().numpy()
    return np.asarray(t)


def cin_pad_of(cin: int) -> int:
    """Channel padding of the packed K axis; the engine derives BK (64/32/16) from it."""
    return (cin + 15) // 16 * 16


def pad_bias(b: np.ndarray) -> np.ndarray:
    n = (b.shape[0] + 255) // 256 * 256
    out = np.zeros(n, np.float32)
    out[: b.shape[0]] = b
    return out


def bn_affine(bn: Optional[Mapping[str, np.ndarray]], cout: int) -> Tuple[np.ndarray, np.ndarray]:
    """y = x * scale + shift for an eval-mode BatchNorm (fp32, as torch computes it)."""
    if bn is None:
        return np.ones(cout, np.float32), np.zeros(cout, np.float32)
    g, b, m, v = (_np(bn[k]).astype(np.float32) for k in ("weight", "bias", "running_mean", "running_var"))
    eps = np.float32(bn.get("eps", 1e-5))
    scale = g / np.sqrt(v + eps)
    return scale.astype(np.float32), (b - m * scale).astype(np.float32)


def pack_conv(weight, bias=None, bn=None) -> Tuple[np.ndarray, np.ndarray]:
    """[Cout,Cin,KH,KW] fp32 (+bias, +BN) -> (fp16 [Cout, KH*KW*Cin_pad], fp32 bias padded to 256)."""
    w = _np(weight).astype(np.float32)
    cout, cin, kh, kw = w.shape
    scale, shift = bn_affine(bn, cout)
    b = np.zeros(cout, np.float32) if bias is None else _np(bias).astype(np.float32)
    w = w * scale[:, None, None, None]
    b = b * scale + shift
    cp = cin_pad_of(cin)
    packed = np.zeros((cout, kh * kw, cp), np.float16)
    packed[:, :, :cin] = w.transpose(0, 2, 3, 1).reshape(cout, kh * kw, cin).astype(np.float16)
    return packed.reshape(cout, kh * kw * cp), pad_bias(b)


def split_packed(packed_f32: np.ndarray, taps: int = 1) -> np.ndarray:
    """fp32 K-major weight matrix [Cout, taps * Cin_pad] -> the split-fp16 operand [Cout, taps * 3 * Cin_pad]: per filter tap
    [W_hi | W_lo | W_hi] with W_hi = fp16(W), W_lo = fp16(W - W_hi).  Against activations stored as [hi | lo] channel pairs the
    k-block walk (hi, hi, lo) of conv_igemm_tcgen05 accumulates A_hi W_hi + A_hi W_lo + A_lo W_hi in fp32: the product to
    ~2^-21 relative at 3x the MMAs (the "fp32x" precision mode)."""
    cout, k = packed_f32.shape
    w = packed_f32.reshape(cout, taps, k // taps).astype(np.float32)
    hi = w.astype(np.float16)
    lo = (w - hi.astype(np.float32)).astype(np.float16)
    return np.concatenate([hi, lo, hi], 2).reshape(cout, taps * 3 * (k // taps))


def conv_matrix_f32(weight, bias=None, bn=None) -> Tuple[np.ndarray, np.ndarray, int]:
    """pack_conv before the fp16 cast: (fp32 [Cout, KH*KW*Cin_pad], fp32 bias padded to 256, taps)."""
    w = _np(weight).astype(np.float32)
    cout, cin, kh, kw = w.shape
    scale, shift = bn_affine(bn, cout)
    b = np.zeros(cout, np.float32) if bias is None else _np(bias).astype(np.float32)
    w = w * scale[:, None, None, None]
    b = b * scale + shift
    cp = cin_pad_of(cin)
    packed = np.zeros((cout, kh * kw, cp), np.float32)
    packed[:, :, :cin] = w.transpose(0, 2, 3, 1).reshape(cout, kh * kw, cin)
    return packed.reshape(cout, kh * kw * cp), pad_bias(b), kh * kw


def pack_conv_split(weight, bias=None, bn=None, flat: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """pack_conv for the fp32x mode.  flat: the conv is run as ONE flat GEMM over a re-laid-out input whose row already holds
    all taps (K = taps * Cin), so the hi / lo / hi triple spans the whole K instead of each tap."""
    w, b, taps = conv_matrix_f32(weight, bias, bn)
    return split_packed(w, 1 if flat else taps), b


def pack_linear(weight, bias=None) -> Tuple[np.ndarray, np.ndarray]:
    """nn.Linear weight [out, in] -> same packing as a 1x1 conv."""
    w = _np(weight).astype(np.float32)
    return pack_conv(w[:, :, None, None], bias)


def pack_stem7x7(weight, bn=None, split: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """7x7 stride-2 stem on a 3-channel image: K index = r*32 + s*4 + c (s padded 7->8, c padded 3->4),
    matching the overlapping-window TMA view of the padded 4-channel input (csrc/igemm_host.cu, A_STEM).
    split: fp32x triples per filter row (``split_packed``)."""
    w = _np(weight).astype(np.float32)
    cout, cin, kh, kw = w.shape
    assert (cin, kh, kw) == (3, 7, 7), w.shape
    scale, shift = bn_affine(bn, cout)
    w = w * scale[:, None, None, None]
    packed = np.zeros((cout, 7, 8, 4), np.float32)
    packed[:, :, :7, :3] = w.transpose(0, 2, 3, 1)
    packed = packed.reshape(cout, 7 * 32)
    return (split_packed(packed, 7) if split else packed.astype(np.float16)), pad_bias(shift)


def pack_deconv2x2(weight, bias=None, bn=None, split: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """ConvTranspose2d(k=2,s=2) weight [Cin,Cout,2,2] -> GEMM weight [(dy*2+dx)*Cout + co][Cin_pad]
    whose output is pixel-shuffled by the epilogue (OUT_SHUF2)."""
    w = _np(weight).astype(np.float32)
    cin, cout, kh, kw = w.shape
    assert (kh, kw) == (2, 2)
    scale, shift = bn_affine(bn, cout)
    b = np.zeros(cout, np.float32) if bias is None else _np(bias).astype(np.float32)
    b = b * scale + shift
    cp = cin_pad_of(cin)
    packed = np.zeros((4, cout, cp), np.float32)
    # w[ci, co, dy, dx] -> [q=dy*2+dx, co, ci]
    packed[:, :, :cin] = (w * scale[None, :, None, None]).transpose(2, 3, 1, 0).reshape(4, cout, cin)
    packed = packed.reshape(4 * cout, cp)
    return (split_packed(packed, 1) if split else packed.astype(np.float16)), pad_bias(np.tile(b, 4))


def write_blob(tensors: Mapping[str, np.ndarray]) -> bytes:
    names = list(tensors.keys())
    entries = []
    payload = bytearray()
    for name in names:
        a = np.ascontiguousarray(tensors[name])
        if a.dtype not in _DT:
            raise TypeError(f"{name}: unsupported dtype {a.dtype}")
        if a.ndim > 4 or a.ndim == 0:
            raise ValueError(f"{name}: rank {a.ndim}")
        while len(payload) % 256:
            payload.append(0)
        off = len(payload)
        payload += a.tobytes()
        dims = list(a.shape) + [0] * (4 - a.ndim)
        nm = name.encode()
        if len(nm) > 95:
            raise ValueError(f"tensor name too long: {name}")
        entries.append(struct.pack("<96sII4IQQ", nm, _DT[a.dtype], a.ndim, *dims, off, a.nbytes))
    header_size = 8 + 4 + 4 + 8 + 8
    table = b"".join(entries)
    data_offset = (header_size + len(table) + 255) // 256 * 256
    head = struct.pack("<8sIIQQ", b"DVWBLOB1", len(names), 0, data_offset, len(payload))
    pad = b"\0" * (data_offset - header_size - len(table))
    return head + table + pad + bytes(payload)


# --------------------------------------------------------------------------- DBNet-R18
def _bn(sd, prefix):
    return {k: _np(sd[f"{prefix}.{k}"]) for k in ("weight", "bias", "running_mean", "running_var")}


def pack_dbnet_r18(sd: Mapping[str, "np.ndarray"], precise: bool = False) -> bytes:
    """state_dict of the reference DBModel (db_net/dbnet.py:715-728; keys backbone.* / decoder.*).  precise=True packs every
    weight as a split-fp16 triple per filter tap and marks the blob with a "precision" entry (fp32x mode, csrc/dbnet.cu)."""
    t: Dict[str, np.ndarray] = {}

    def put(name, wb):
        t[name + ".w"], t[name + ".b"] = wb

    pack_conv = pack_conv_split if precise else globals()["pack_conv"]
    if precise:
        t["precision"] = np.array([1], np.int32)
    put("stem", pack_stem7x7(sd["backbone.conv1.weight"], _bn(sd, "backbone.bn1"), split=precise))
    for L in range(1, 5):
        for B in range(2):
            p = f"backbone.layer{L}.{B}"
            put(f"layer{L}.{B}.conv1", pack_conv(sd[f"{p}.conv1.weight"], None, _bn(sd, f"{p}.bn1")))
            put(f"layer{L}.{B}.conv2", pack_conv(sd[f"{p}.conv2.weight"], None, _bn(sd, f"{p}.bn2")))
            if f"{p}.downsample.0.weight" in sd:
                put(f"layer{L}.{B}.down", pack_conv(sd[f"{p}.downsample.0.weight"], None, _bn(sd, f"{p}.downsample.1")))
    for n in ("in5", "in4", "in3", "in2"):
        put(n, pack_conv(sd[f"decoder.{n}.weight"], sd.get(f"decoder.{n}.bias")))
    for n in ("out5", "out4", "out3"):
        put(n, pack_conv(sd[f"decoder.{n}.0.weight"], sd.get(f"decoder.{n}.0.bias")))
    put("out2", pack_conv(sd["decoder.out2.weight"], sd.get("decoder.out2.bias")))
    put("bin.conv", pack_conv(sd["decoder.binarize.0.weight"], sd.get("decoder.binarize.0.bias"), _bn(sd, "decoder.binarize.1")))
    put("bin.deconv1", pack_deconv2x2(sd["decoder.binarize.3.weight"], sd["decoder.binarize.3.bias"], _bn(sd, "decoder.binarize.4"), split=precise))
    w2 = _np(sd["decoder.binarize.6.weight"]).astype(np.float32)  # [64,1,2,2]
    t["bin.deconv2.w"] = w2.reshape(64, 4).astype(np.float16)
    if precise:
        t["bin.deconv2.w32"] = np.ascontiguousarray(w2.reshape(64, 4))
    t["bin.deconv2.b"] = _np(sd["decoder.binarize.6.bias"]).astype(np.float32).reshape(1)
    return write_blob(t)


# --------------------------------------------------------------------------- ConvNextViT
def pack_convnext_vit(sd: Mapping[str, "np.ndarray"], precise: bool = False) -> bytes:
    """state_dict of the reference ConvNextViT (convnext_vit/modeling_convnext_vit.py:20-45).  precise=True packs every GEMM
    weight as a split-fp16 triple (``split_packed``) and marks the blob with a "precision" entry: the engine then runs its
    fp32x mode (csrc/convnextvit.cu).

    layer_scale is folded into pwconv2 (gamma * (W x + b)), the attention scale 1/sqrt(64) (an exact power of
    two) into the query projection, q/k/v are concatenated into one [576,192] GEMM, the (2,1) down-sampling conv
    becomes a flat [C', 2C] GEMM over the re-laid-out LayerNorm output, and position_embeddings[:, 1:] is kept as
    an fp32 [75,192] table added by the patch-projection epilogue."""
    t: Dict[str, np.ndarray] = {}
    f = lambda k: _np(sd[k]).astype(np.float32)

    def put(name, wb):
        t[name + ".w"], t[name + ".b"] = wb

    def lin(weight, bias=None):
        w = _np(weight).astype(np.float32)
        return pack_conv_split(w[:, :, None, None], bias) if precise else pack_linear(w, bias)

    def conv_flat(weight, bias=None):  # (2,1) / 1x1 convs that the engine runs as one flat GEMM
        return pack_conv_split(weight, bias, flat=True) if precise else pack_conv(weight, bias)

    if precise:
        t["precision"] = np.array([1], np.int32)

    p = "cnn_model.embeddings"
    t["patch.w"] = np.ascontiguousarray(f(p + ".patch_embeddings.weight").reshape(96, 16).T)  # [16][96], k = dy*4+dx
    t["patch.b"] = f(p + ".patch_embeddings.bias")
    t["patch.ln.w"], t["patch.ln.b"] = f(p + ".layernorm.weight"), f(p + ".layernorm.bias")
    blk = 0
    depths, dims = (3, 3, 8, 3), (96, 192, 256, 512)
    for s, (depth, dim) in enumerate(zip(depths, dims)):
        sp = f"cnn_model.encoder.stages.{s}"
        if s > 0:
            t[f"ds{s}.ln.w"], t[f"ds{s}.ln.b"] = f(sp + ".downsampling_layer.0.weight"), f(sp + ".downsampling_layer.0.bias")
            put(f"ds{s}.conv", conv_flat(f(sp + ".downsampling_layer.1.weight"), f(sp + ".downsampling_layer.1.bias")))
        for j in range(depth):
            lp = f"{sp}.layers.{j}"
            gamma = f(lp + ".layer_scale_parameter")
            t[f"blk{blk}.dw.w"] = np.ascontiguousarray(f(lp + ".dwconv.weight").reshape(dim, 49).T)  # [49][C]
            t[f"blk{blk}.dw.b"] = f(lp + ".dwconv.bias")
            t[f"blk{blk}.ln.w"], t[f"blk{blk}.ln.b"] = f(lp + ".layernorm.weight"), f(lp + ".layernorm.bias")
            put(f"blk{blk}.pw1", lin(f(lp + ".pwconv1.weight"), f(lp + ".pwconv1.bias")))
            put(f"blk{blk}.pw2", lin(f(lp + ".pwconv2.weight") * gamma[:, None], f(lp + ".pwconv2.bias") * gamma))
            blk += 1
    v = "vitstr.vit"
    put("vit.proj", conv_flat(f(v + ".embeddings.patch_embeddings.projection.weight"),
                              f(v + ".embeddings.patch_embeddings.projection.bias")))
    t["vit.pos"] = np.ascontiguousarray(f(v + ".embeddings.position_embeddings")[0, 1:, :])
    L = 0
    while f"{v}.encoder.layer.{L}.attention.attention.query.weight" in sd:
        lp = f"{v}.encoder.layer.{L}"
        a = lp + ".attention.attention"
        wq, bq = f(a + ".query.weight") * np.float32(0.125), f(a + ".query.bias") * np.float32(0.125)
        put(f"vit{L}.qkv", lin(np.concatenate([wq, f(a + ".key.weight"), f(a + ".value.weight")], 0),
                                       np.concatenate([bq, f(a + ".key.bias"), f(a + ".value.bias")], 0)))
        put(f"vit{L}.proj", lin(f(lp + ".attention.output.dense.weight"), f(lp + ".attention.output.dense.bias")))
        put(f"vit{L}.fc1", lin(f(lp + ".intermediate.dense.weight"), f(lp + ".intermediate.dense.bias")))
        put(f"vit{L}.fc2", lin(f(lp + ".output.dense.weight"), f(lp + ".output.dense.bias")))
        t[f"vit{L}.ln1.w"], t[f"vit{L}.ln1.b"] = f(lp + ".layernorm_before.weight"), f(lp + ".layernorm_before.bias")
        t[f"vit{L}.ln2.w"], t[f"vit{L}.ln2.b"] = f(lp + ".layernorm_after.weight"), f(lp + ".layernorm_after.bias")
        L += 1
    t["vit.ln.w"], t["vit.ln.b"] = f(v + ".layernorm.weight"), f(v + ".layernorm.bias")
    put("cls", lin(f("vitstr.classifier.weight"), f("vitstr.classifier.bias")))
    return write_blob(t)


# --------------------------------------------------------------------------- Lore (DLA-34 + DCNv2 detector)
def pack_stem7x7_s1(weight, bn=None, split: bool = False) -> Tuple[np.ndarray, np.ndarray]:
    """7x7 stride-1 stem on a 3-channel image: K index = r*64 + s*8 + c (s padded 7->8, c padded 3->8), matching the
    overlapping-window TMA view of the zero-bordered 8-channel input (csrc/igemm_host.cu, A_STEM stride 1)."""
    w = _np(weight).astype(np.float32)
    cout, cin, kh, kw = w.shape
    assert (cin, kh, kw) == (3, 7, 7), w.shape
    scale, shift = bn_affine(bn, cout)
    w = w * scale[:, None, None, None]
    packed = np.zeros((cout, 7, 8, 8), np.float32)
    packed[:, :, :7, :3] = w.transpose(0, 2, 3, 1)
    packed = packed.reshape(cout, 7 * 64)
    return (split_packed(packed, 7) if split else packed.astype(np.float16)), pad_bias(shift)


def pack_win3x3_c16(weight, bn=None) -> Tuple[np.ndarray, np.ndarray]:
    """3x3 stride-1 conv on 16 channels for conv_win_tcgen05: K index = r*64 + s*16 + c with the window padded from 3 to 4
    pixels (s = 3 carries zero weights), i.e. one k-block per filter row over a 128-byte window of the zero-bordered input."""
    w = _np(weight).astype(np.float32)
    cout, cin, kh, kw = w.shape
    assert (cin, kh, kw) == (16, 3, 3), w.shape
    scale, shift = bn_affine(bn, cout)
    w = w * scale[:, None, None, None]
    packed = np.zeros((cout, 3, 4, 16), np.float16)
    packed[:, :, :3, :] = w.transpose(0, 2, 3, 1).astype(np.float16)
    return packed.reshape(cout, 3 * 64), pad_bias(shift)


def pack_win3x3_c16_planar(weight, bn=None) -> Tuple[np.ndarray, np.ndarray]:
    """As pack_win3x3_c16 for the PATCH mode of conv_win_tcgen05, whose staged patch is two planes of 8 channels:
    K index = r*64 + plane*32 + s*8 + c8."""
    w = _np(weight).astype(np.float32)
    cout, cin, kh, kw = w.shape
    assert (cin, kh, kw) == (16, 3, 3), w.shape
    scale, shift = bn_affine(bn, cout)
    w = (w * scale[:, None, None, None]).transpose(0, 2, 3, 1)  # [cout, r, s, c]
    packed = np.zeros((cout, 3, 2, 4, 8), np.float16)
    for plane in range(2):
        packed[:, :, plane, :3, :] = w[:, :, :, plane * 8:(plane + 1) * 8].astype(np.float16)
    return packed.reshape(cout, 3 * 64), pad_bias(shift)


def pack_split_linear(weight, bias=None) -> Tuple[np.ndarray, np.ndarray]:
    """nn.Linear weight [out, in] -> split-fp16 operand [out, 3*in_pad] = [W_hi | W_lo | W_hi] with W_hi = fp16(W),
    W_lo = fp16(W - W_hi): against activations stored as [hi | lo] the k-block walk (hi, hi, lo) accumulates
    A_hi W_hi + A_hi W_lo + A_lo W_hi in fp32, i.e. the product to ~2^-21 relative (csrc/igemm_host.cu plan_linear)."""
    w = _np(weight).astype(np.float32)
    cout, cin = w.shape
    cp = (cin + 63) // 64 * 64 if cin >= 64 else cin_pad_of(cin)
    hi = np.zeros((cout, cp), np.float16)
    lo = np.zeros((cout, cp), np.float16)
    hi[:, :cin] = w.astype(np.float16)
    lo[:, :cin] = (w - hi[:, :cin].astype(np.float32)).astype(np.float16)
    b = np.zeros(cout, np.float32) if bias is None else _np(bias).astype(np.float32)
    return np.concatenate([hi, lo, hi], 1), pad_bias(b)


def _pad_rows(w: np.ndarray, b: np.ndarray, rows: int) -> Tuple[np.ndarray, np.ndarray]:
    out = np.zeros((rows, w.shape[1]), w.dtype)
    out[: w.shape[0]] = w
    return out, b


LORE_SMALL_HEADS = (("hm", 2), ("reg", 2), ("wh", 8), ("st", 8))  # channel order of the packed 24-wide map


def pack_lore_dla34(sd: Mapping[str, "np.ndarray"], precise: bool = False) -> bytes:
    """state_dict of the reference `get_dla_dcn(34, heads)` (lore/lore_dla_34.py:193-206) -> engine blob.  precise=True packs
    every conv weight as a split-fp16 triple (per filter tap; over the whole K for the convs that run as flat GEMMs over gathered
    columns: DCN main conv, ax / cr 3x3) and marks the blob with a "precision" entry (fp32x mode, csrc/lore_net.cu).

    BatchNorm is folded everywhere (incl. the BN after each DCN).  conv_offset_mask keeps its 27 outputs padded to 32
    (fp32 epilogue).  The four small heads share one 3x3 conv (64 -> 4*256) and one block-diagonal 1x1 (1024 -> 24);
    `ax` / `cr` keep separate weights because they are evaluated only at decoded points."""
    t: Dict[str, np.ndarray] = {}
    f = lambda k: _np(sd[k]).astype(np.float32)

    def put(name, wb):
        t[name + ".w"], t[name + ".b"] = wb

    pack_conv = pack_conv_split if precise else globals()["pack_conv"]
    flat = (lambda w, b=None, bn=None: pack_conv_split(w, b, bn, flat=True)) if precise else globals()["pack_conv"]
    if precise:
        t["precision"] = np.array([1], np.int32)
    put("base", pack_stem7x7_s1(sd["base.base_layer.0.weight"], _bn(sd, "base.base_layer.1"), split=precise))
    put("level0", pack_conv(sd["base.level0.0.weight"], None, _bn(sd, "base.level0.1")))
    if not precise:  # the window-conv kernel keeps fp16 weights in shared memory: fp32x runs these two layers on conv_igemm_tcgen05
        put("level0.win", pack_win3x3_c16(sd["base.level0.0.weight"], _bn(sd, "base.level0.1")))
        put("level0.winp", pack_win3x3_c16_planar(sd["base.level0.0.weight"], _bn(sd, "base.level0.1")))
    put("level1", pack_conv(sd["base.level1.0.weight"], None, _bn(sd, "base.level1.1")))
    if not precise:  # stride-2 window form for conv_win_tcgen05 (the same 4-pixel x 16-channel window as level0.win)
        put("level1.win", pack_win3x3_c16(sd["base.level1.0.weight"], _bn(sd, "base.level1.1")))
    for k in sd:
        if not k.startswith("base.level") or k.startswith(("base.level0", "base.level1")):
            continue
        if k.endswith(".conv1.weight") or k.endswith(".conv2.weight"):
            p, n = k[: -len(".weight")].rsplit(".", 1)
            put(k[5: -len(".weight")], pack_conv(sd[k], None, _bn(sd, f"{p}.bn{n[-1]}")))
        elif k.endswith(".root.conv.weight"):
            p = k[: -len(".conv.weight")]
            put(k[5: -len(".conv.weight")], pack_conv(sd[k], None, _bn(sd, p + ".bn")))
        elif k.endswith(".project.0.weight"):
            p = k[: -len(".0.weight")]
            put(k[5: -len(".0.weight")], pack_conv(sd[k], None, _bn(sd, p + ".1")))
    for k in sd:
        if k.endswith(".conv.conv_offset_mask.weight"):
            p = k[: -len(".conv.conv_offset_mask.weight")]  # e.g. dla_up.ida_0.proj_1
            put(p + ".dcn", flat(sd[p + ".conv.weight"], sd[p + ".conv.bias"], _bn(sd, p + ".actf.0")))
            w, b = pack_conv(sd[k], sd[p + ".conv.conv_offset_mask.bias"])
            put(p + ".om", _pad_rows(w, b, 32))
        elif ".up_" in k and k.endswith(".weight"):
            w = f(k)  # [C,1,2f,2f] -> fp32 [2f][2f][C]
            t[k[: -len(".weight")] + ".w"] = np.ascontiguousarray(w[:, 0].transpose(1, 2, 0))
    # small heads: one 3x3 conv and one block-diagonal 1x1
    w3 = np.concatenate([f(f"{h}.0.weight") for h, _ in LORE_SMALL_HEADS], 0)
    b3 = np.concatenate([f(f"{h}.0.bias") for h, _ in LORE_SMALL_HEADS], 0)
    put("heads.conv", pack_conv(w3, b3))
    w1 = np.zeros((24, 1024, 1, 1), np.float32)
    b1 = np.zeros(24, np.float32)
    row = 0
    for i, (h, c) in enumerate(LORE_SMALL_HEADS):
        w1[row: row + c, 256 * i: 256 * (i + 1)] = f(f"{h}.2.weight")
        b1[row: row + c] = f(f"{h}.2.bias")
        row += c
    put("heads.out", pack_conv(w1, b1))
    for h in ("ax", "cr"):
        put(f"{h}.conv", flat(f(f"{h}.0.weight"), f(f"{h}.0.bias")))
        put(f"{h}.out", pack_conv(f(f"{h}.2.weight"), f(f"{h}.2.bias")))
    return write_blob(t)


def deconv4x4_as_conv3x3(weight, bn=None) -> Tuple[np.ndarray, np.ndarray]:
    """ConvTranspose2d(k=4, s=2, p=1, no bias) weight [Cin, Cout, 4, 4] (+BN) -> the equivalent 3x3 stride-1 pad-1 conv weight
    [4*Cout, Cin, 3, 3] whose output is pixel-shuffled (OUT_SHUF2, row (py*2 + px)*Cout + co) plus its bias [4*Cout].
    Output pixel (2y + py, 2x + px) receives input (y + dy, x + dx) through kernel tap ky = py + 1 - 2 dy (same in x): two of the
    three conv rows per parity, the third stays zero."""
    w = _np(weight).astype(np.float32)
    cin, cout, kh, kw = w.shape
    assert (kh, kw) == (4, 4), w.shape
    scale, shift = bn_affine(bn, cout)
    ws = w * scale[None, :, None, None]
    out = np.zeros((2, 2, cout, cin, 3, 3), np.float32)
    for py in range(2):
        for r in range(3):
            ky = py + 3 - 2 * r
            if not 0 <= ky < 4:
                continue
            for px in range(2):
                for c in range(3):
                    kx = px + 3 - 2 * c
                    if 0 <= kx < 4:
                        out[py, px, :, :, r, c] = ws[:, :, ky, kx].T
    return out.reshape(4 * cout, cin, 3, 3), np.tile(shift, 4).astype(np.float32)


LORE_R18_HEAD_ORDER = ("hm", "reg", "wh", "st", "ax", "cr")  # channel-slice order of the 384-wide head buffers


def pack_lore_resnet18(sd: Mapping[str, "np.ndarray"], precise: bool = False) -> bytes:
    """state_dict of the reference `LoreDetectModel` (lore/lore_detector.py:148-389, the `wireless` configuration) -> engine blob
    (model kind "lore_resnet18", csrc/lore_net.cu build_r18).  BatchNorm folded everywhere; each ConvTranspose 4x4 s2 becomes a
    3x3 conv to 4 x 256 pixel-shuffled channels (deconv4x4_as_conv3x3); the first conv of the six heads is ONE 3x3 conv
    256 -> 6 x 64 (order LORE_R18_HEAD_ORDER), the 64 -> 64 convs stay per head, the last 1x1 of hm / reg / wh / st is one
    block-diagonal 256 -> 24 conv (LORE_SMALL_HEADS columns), `ax` / `cr` keep their 64 -> 256 matrices for the sparse
    evaluation at decoded points."""
    t: Dict[str, np.ndarray] = {}
    f = lambda k: _np(sd[k]).astype(np.float32)

    def put(name, wb):
        t[name + ".w"], t[name + ".b"] = wb

    pack = pack_conv_split if precise else pack_conv
    if precise:
        t["precision"] = np.array([1], np.int32)
    put("stem", pack_stem7x7(sd["conv1.weight"], _bn(sd, "bn1"), split=precise))
    for L in range(1, 5):
        for B in range(2):
            p = f"layer{L}.{B}"
            put(p + ".conv1", pack(sd[p + ".conv1.weight"], sd[p + ".conv1.bias"], _bn(sd, p + ".bn1")))
            put(p + ".conv2", pack(sd[p + ".conv2.weight"], sd[p + ".conv2.bias"], _bn(sd, p + ".bn2")))
            if (p + ".downsample.0.weight") in sd:
                put(p + ".down", pack(sd[p + ".downsample.0.weight"], None, _bn(sd, p + ".downsample.1")))
    for name in ("adaption3", "adaption2", "adaption1", "adaption0", "adaptionU1"):
        put(name, pack(sd[name + ".weight"]))
    for i in range(1, 5):
        w3, b3 = deconv4x4_as_conv3x3(sd[f"deconv_layers{i}.0.weight"], _bn(sd, f"deconv_layers{i}.1"))
        put(f"up{i}", pack(w3, b3))
    hc = 64
    put("heads.conv1", pack(np.concatenate([f(f"{h}.0.weight") for h in LORE_R18_HEAD_ORDER], 0),
                            np.concatenate([f(f"{h}.0.bias") for h in LORE_R18_HEAD_ORDER], 0)))
    for h in LORE_R18_HEAD_ORDER:
        if h == "reg":
            continue
        for j in (2, 4, 6):
            put(f"heads.{h}.{j}", pack(f(f"{h}.{j}.weight"), f(f"{h}.{j}.bias")))
    w1 = np.zeros((24, 4 * hc, 1, 1), np.float32)
    b1 = np.zeros(24, np.float32)
    row = 0
    for i, (h, c) in enumerate(LORE_SMALL_HEADS):
        assert LORE_R18_HEAD_ORDER[i] == h
        last = 2 if h == "reg" else 8
        w1[row: row + c, hc * i: hc * (i + 1)] = f(f"{h}.{last}.weight")
        b1[row: row + c] = f(f"{h}.{last}.bias")
        row += c
    put("heads.out", pack(w1, b1))
    for h in ("ax", "cr"):
        put(f"{h}.out", pack(f(f"{h}.8.weight"), f(f"{h}.8.bias")))
    return write_blob(t)


def pack_lore_processor(sd: Mapping[str, "np.ndarray"]) -> bytes:
    """state_dict of the reference LoreProcessModel (lore/lore_processor.py:399-514) -> engine blob.  Every Linear is
    packed for the split-fp16 GEMM (the cell counts are tiny, the outputs are ROUNDED to integers downstream, so this
    model runs at ~fp32 accuracy); q/k/v are concatenated, 1/sqrt(d_k) is folded into q."""
    t: Dict[str, np.ndarray] = {}
    f = lambda k: _np(sd[k]).astype(np.float32)

    def put(name, w, b):
        t[name + ".w"], t[name + ".b"] = pack_split_linear(w, b)

    def transformer(src, dst):
        put(dst + ".in", f(src + ".linear.weight"), f(src + ".linear.bias"))
        L = 0
        while f"{src}.encoder.layers.{L}.norm_1.alpha" in sd:
            lp, dp = f"{src}.encoder.layers.{L}", f"{dst}.{L}"
            for n in ("norm_1", "norm_2"):
                t[f"{dp}.{n}.a"], t[f"{dp}.{n}.b"] = f(f"{lp}.{n}.alpha"), f(f"{lp}.{n}.bias")
            sc = np.float32(1.0 / np.sqrt(32.0))
            wq, bq = f(lp + ".attn.q_linear.weight") * sc, f(lp + ".attn.q_linear.bias") * sc
            put(dp + ".qkv", np.concatenate([wq, f(lp + ".attn.k_linear.weight"), f(lp + ".attn.v_linear.weight")], 0),
                np.concatenate([bq, f(lp + ".attn.k_linear.bias"), f(lp + ".attn.v_linear.bias")], 0))
            put(dp + ".out", f(lp + ".attn.out.weight"), f(lp + ".attn.out.bias"))
            put(dp + ".ff1", f(lp + ".ff.linear_1.weight"), f(lp + ".ff.linear_1.bias"))
            put(dp + ".ff2", f(lp + ".ff.linear_2.weight"), f(lp + ".ff.linear_2.bias"))
            L += 1
        put(dst + ".dec0", f(src + ".decoder.linear.0.weight"), f(src + ".decoder.linear.0.bias"))
        put(dst + ".dec2", f(src + ".decoder.linear.2.weight"), f(src + ".decoder.linear.2.bias"))
        return L

    n_axis = transformer("tsfm_axis", "axis")
    n_stack = transformer("stacker.tsfm", "stack")
    put("stack.enc0", f("stacker.logi_encoder.0.weight"), f("stacker.logi_encoder.0.bias"))
    put("stack.enc2", f("stacker.logi_encoder.2.weight"), f("stacker.logi_encoder.2.bias"))
    t["x_pos"], t["y_pos"] = f("x_position_embeddings.weight"), f("y_position_embeddings.weight")
    t["meta"] = np.array([n_axis, n_stack, 256, 8], np.int32)
    return write_blob(t)


# --------------------------------------------------------------------------- CenterNet (DLA-34, plain IDA-up)
def pack_centernet_dla34(sd: Mapping[str, "np.ndarray"]) -> bytes:
    """state_dict of the reference CenterNet `DLASeg()` (center_net/modeling_centernet.py:601-668) -> engine blob.  The base
    uses the Lore packing; IDAUp proj / node convs are folded with their BatchNorm; the four heads share one 3x3 conv and one
    block-diagonal 1x1 whose 24 output channels are ordered hm, reg, c2v, v2c (the packed map dv_centernet_decode reads)."""
    t: Dict[str, np.ndarray] = {}
    f = lambda k: _np(sd[k]).astype(np.float32)

    def put(name, wb):
        t[name + ".w"], t[name + ".b"] = wb

    put("base", pack_stem7x7_s1(sd["base.base_layer.0.weight"], _bn(sd, "base.base_layer.1")))
    put("level0", pack_conv(sd["base.level0.0.weight"], None, _bn(sd, "base.level0.1")))
    put("level0.win", pack_win3x3_c16(sd["base.level0.0.weight"], _bn(sd, "base.level0.1")))
    put("level0.winp", pack_win3x3_c16_planar(sd["base.level0.0.weight"], _bn(sd, "base.level0.1")))
    put("level1", pack_conv(sd["base.level1.0.weight"], None, _bn(sd, "base.level1.1")))
    for k in sd:
        if k.startswith("base.level") and not k.startswith(("base.level0", "base.level1")):
            if k.endswith(".conv1.weight") or k.endswith(".conv2.weight"):
                p, n = k[: -len(".weight")].rsplit(".", 1)
                put(k[5: -len(".weight")], pack_conv(sd[k], None, _bn(sd, f"{p}.bn{n[-1]}")))
            elif k.endswith(".root.conv.weight"):
                put(k[5: -len(".conv.weight")], pack_conv(sd[k], None, _bn(sd, k[: -len(".conv.weight")] + ".bn")))
            elif k.endswith(".project.0.weight"):
                put(k[5: -len(".0.weight")], pack_conv(sd[k], None, _bn(sd, k[: -len(".0.weight")] + ".1")))
        elif k.startswith("dla_up.") and k.endswith(".0.weight"):  # proj_k / node_k: conv + BN + ReLU
            p = k[: -len(".0.weight")]
            put(p, pack_conv(sd[k], None, _bn(sd, p + ".1")))
        elif k.startswith("dla_up.") and ".up_" in k and k.endswith(".weight"):
            t[k[: -len(".weight")] + ".w"] = np.ascontiguousarray(f(k)[:, 0].transpose(1, 2, 0))
    order = (("hm", 2), ("reg", 2), ("c2v", 8), ("v2c", 8))
    put("heads.conv", pack_conv(np.concatenate([f(f"{h}.0.weight") for h, _ in order], 0), np.concatenate([f(f"{h}.0.bias") for h, _ in order], 0)))
    w1 = np.zeros((24, 1024, 1, 1), np.float32)
    b1 = np.zeros(24, np.float32)
    row = 0
    for i, (h, c) in enumerate(order):
        w1[row: row + c, 256 * i: 256 * (i + 1)] = f(f"{h}.2.weight")
        b1[row: row + c] = f(f"{h}.2.bias")
        row += c
    put("heads.out", pack_conv(w1, b1))
    return write_blob(t)


# --------------------------------------------------------------------------- CRNN (crnn/modeling_crnn.py)
def pack_crnn(sd: Mapping[str, "np.ndarray"]) -> bytes:
    """state_dict of the reference CRNN module (crnn/modeling_crnn.py:36-88) -> engine blob (model kind "crnn", csrc/crnn.cu).
    BatchNorm folded into every conv; conv0's single gray input channel becomes channel 0 of an 8-channel pixel; conv4's (2,1)
    kernel is packed tap-major (row 0, row 1); per LSTM layer the input projections of both directions are ONE matrix
    [W_ih ; W_ih_reverse] with bias b_ih + b_hh, the recurrent matrices stay per direction (no bias); cls has no bias."""
    t: Dict[str, np.ndarray] = {}

    def put(name, wb):
        t[name + ".w"], t[name + ".b"] = wb

    w0 = _np(sd["conv0.0.weight"]).astype(np.float32)
    w0p = np.zeros((w0.shape[0], 8, 3, 3), np.float32)
    w0p[:, :1] = w0
    put("conv0", pack_conv(w0p, sd["conv0.0.bias"], _bn(sd, "conv0.1")))
    for name, c, b in (("conv1", "conv1.0", "conv1.1"), ("conv2a", "conv2.0", "conv2.1"), ("conv2b", "conv2.3", "conv2.4"),
                       ("conv3a", "conv3.0", "conv3.1"), ("conv3b", "conv3.3", "conv3.4"), ("conv4", "conv4.0", "conv4.1")):
        put(name, pack_conv(sd[c + ".weight"], sd[c + ".bias"], _bn(sd, b)))
    for layer in (0, 1):
        p = f"rnn.{layer}.rnn."
        wih = np.concatenate([_np(sd[p + "weight_ih_l0"]), _np(sd[p + "weight_ih_l0_reverse"])], 0).astype(np.float32)
        bias = np.concatenate([_np(sd[p + "bias_ih_l0"]) + _np(sd[p + "bias_hh_l0"]),
                               _np(sd[p + "bias_ih_l0_reverse"]) + _np(sd[p + "bias_hh_l0_reverse"])]).astype(np.float32)
        put(f"rnn.{layer}.ih", pack_linear(wih, bias))
        t[f"rnn.{layer}.hh.w"] = pack_linear(sd[p + "weight_hh_l0"])[0]
        t[f"rnn.{layer}.hh_reverse.w"] = pack_linear(sd[p + "weight_hh_l0_reverse"])[0]
        put(f"rnn.{layer}.emb", pack_linear(sd[f"rnn.{layer}.embedding.weight"], sd[f"rnn.{layer}.embedding.bias"]))
    t["cls.w"] = pack_linear(sd["cls.weight"])[0]
    return write_blob(t)
