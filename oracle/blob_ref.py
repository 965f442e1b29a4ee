"""Read a packed weight blob back and run DBNet-R18 FROM THE PACKED TENSORS on CPU -- TEST INFRASTRUCTURE.

`read_blob` parses the container that pdf_table_b200/weights.write_blob writes and csrc/capi.cu load_blob reads ("DVWBLOB1"
header, 136-byte entries, 256-byte aligned payload).  `dbnet_r18_from_blob` undoes the packing (K-major fp16 [Cout][tap][Cin_pad]
with BatchNorm folded, the 7x7 stem as [Cout][7][8][4], the 2x2 transposed conv as four pixel-shuffled GEMM blocks) and runs the
network with plain PyTorch ops in the order csrc/dbnet.cu plans them, so the HOST side of the weight path -- folding, padding,
tap order, serialisation -- is checked on CPU against the oracle forward (oracle/dbnet_ref.py, pinned against the reference
DBModel).  The remaining difference is the fp16 rounding of the packed weights.
"""
from __future__ import annotations

import struct
from typing import Dict

import numpy as np
import torch
import torch.nn.functional as F

_DT = {0: np.float32, 1: np.float16, 2: np.int32}


def read_blob(blob: bytes, dtypes=None) -> Dict[str, np.ndarray]:
    dtypes = _DT if dtypes is None else dtypes
    magic, n, _, data_offset, data_bytes = struct.unpack_from("<8sIIQQ", blob, 0)
    if magic != b"DVWBLOB1":
        raise ValueError("bad blob magic")
    if data_offset + data_bytes > len(blob):
        raise ValueError("blob truncated")
    out = {}
    for i in range(n):
        name, dt, ndim, d0, d1, d2, d3, off, nbytes = struct.unpack_from("<96sII4IQQ", blob, 32 + i * 136)
        if off % 256 or off + nbytes > data_bytes:
            raise ValueError("blob entry out of range / misaligned")
        shape = (d0, d1, d2, d3)[:ndim]
        a = np.frombuffer(blob, dtype=dtypes[dt], count=int(np.prod(shape)), offset=data_offset + off).reshape(shape)
        out[name.rstrip(b"\0").decode()] = a
    return out


def _unpack_conv(t, name: str, cin: int, k: int):
    w = torch.from_numpy(t[name + ".w"].astype(np.float32))
    cout = w.shape[0]
    w = w.reshape(cout, k * k, -1)[:, :, :cin].reshape(cout, k, k, cin).permute(0, 3, 1, 2).contiguous()
    return w, torch.from_numpy(t[name + ".b"][:cout].copy())


@torch.no_grad()
def dbnet_r18_from_blob(t: Dict[str, np.ndarray], x: torch.Tensor) -> torch.Tensor:
    """t = read_blob(pack_dbnet_r18(sd)); x fp32 [N,3,H,W] -> probability map fp32 [N,1,H,W]."""
    def conv(name, inp, cin, k, stride=1, relu=False):
        w, b = _unpack_conv(t, name, cin, k)
        y = F.conv2d(inp, w, b, stride=stride, padding=k // 2)
        return F.relu(y) if relu else y

    ws = torch.from_numpy(t["stem.w"].astype(np.float32)).reshape(64, 7, 8, 4)[:, :, :7, :3].permute(0, 3, 1, 2).contiguous()
    y = F.relu(F.conv2d(x.float(), ws, torch.from_numpy(t["stem.b"][:64].copy()), stride=2, padding=3))
    y = F.max_pool2d(y, 3, 2, 1)
    feats, cin = [], 64
    for L, cout in zip(range(1, 5), (64, 128, 256, 512)):
        for B in range(2):
            stride = 2 if (L > 1 and B == 0) else 1
            p = f"layer{L}.{B}"
            o = conv(p + ".conv1", y, cin, 3, stride, relu=True)
            o = conv(p + ".conv2", o, cout, 3)
            if (p + ".down.w") in t:
                y = conv(p + ".down", y, cin, 1, stride)
            y = F.relu(o + y)
            cin = cout
        feats.append(y)
    c2, c3, c4, c5 = feats
    in5, in4, in3, in2 = conv("in5", c5, 512, 1), conv("in4", c4, 256, 1), conv("in3", c3, 128, 1), conv("in2", c2, 64, 1)
    up = lambda v, s: F.interpolate(v, scale_factor=s, mode="nearest")
    out4 = up(in5, 2) + in4
    out3 = up(out4, 2) + in3
    out2 = up(out3, 2) + in2
    fuse = torch.cat((up(conv("out5", in5, 256, 3), 8), up(conv("out4", out4, 256, 3), 4), up(conv("out3", out3, 256, 3), 2),
                      conv("out2", out2, 256, 3)), 1)
    b = conv("bin.conv", fuse, 256, 3, relu=True)
    # bin.deconv1: GEMM rows (dy*2+dx)*64 + co over 64 input channels, pixel-shuffled; BN folded, then ReLU
    wd = torch.from_numpy(t["bin.deconv1.w"].astype(np.float32))[:, :64].reshape(2, 2, 64, 64)  # [dy][dx][co][ci]
    bd = torch.from_numpy(t["bin.deconv1.b"][:64].copy())
    b = F.relu(F.conv_transpose2d(b, wd.permute(3, 2, 0, 1).contiguous(), bd, stride=2))
    w2 = torch.from_numpy(t["bin.deconv2.w"].astype(np.float32)).reshape(64, 1, 2, 2)
    b = F.conv_transpose2d(b, w2, torch.from_numpy(t["bin.deconv2.b"].copy()), stride=2)
    return torch.sigmoid(b)


def _unpack_linear(t, name: str, cin: int):
    w = t[name + ".w"].astype(np.float32)
    return w[:, :cin].copy(), t[name + ".b"][:w.shape[0]].copy()


def convnext_vit_state_dict_from_blob(t: Dict[str, np.ndarray]) -> Dict[str, np.ndarray]:
    """t = read_blob(pack_convnext_vit(sd)) -> a state_dict with the REFERENCE's keys that computes what the packed tensors
    compute: layer_scale folded into pwconv2 (so layer_scale_parameter = 1 here), the 1/8 attention scale folded into the
    query projection (undone: x 8, exact), q | k | v split again, the (2,1) down-sampling conv and the patch projection back
    in conv layout, position embeddings with a zero CLS row.  Feeding it to oracle/convnextvit_ref.convnextvit_forward checks
    the packer on CPU; the only difference from the original state_dict is the fp16 rounding of the GEMM weights."""
    sd: Dict[str, np.ndarray] = {}
    p = "cnn_model.embeddings"
    sd[p + ".patch_embeddings.weight"] = np.ascontiguousarray(t["patch.w"].T).reshape(96, 1, 4, 4)
    sd[p + ".patch_embeddings.bias"] = t["patch.b"].copy()
    sd[p + ".layernorm.weight"], sd[p + ".layernorm.bias"] = t["patch.ln.w"].copy(), t["patch.ln.b"].copy()
    blk = 0
    dims = (96, 192, 256, 512)
    for s, (depth, dim) in enumerate(zip((3, 3, 8, 3), dims)):
        sp = f"cnn_model.encoder.stages.{s}"
        if s > 0:
            cin = dims[s - 1]
            sd[sp + ".downsampling_layer.0.weight"], sd[sp + ".downsampling_layer.0.bias"] = t[f"ds{s}.ln.w"].copy(), t[f"ds{s}.ln.b"].copy()
            w = t[f"ds{s}.conv.w"].astype(np.float32).reshape(dim, 2, -1)[:, :, :cin]  # [C'][tap = kh][C]
            sd[sp + ".downsampling_layer.1.weight"] = np.ascontiguousarray(w.transpose(0, 2, 1)).reshape(dim, cin, 2, 1)
            sd[sp + ".downsampling_layer.1.bias"] = t[f"ds{s}.conv.b"][:dim].copy()
        for j in range(depth):
            lp = f"{sp}.layers.{j}"
            sd[lp + ".dwconv.weight"] = np.ascontiguousarray(t[f"blk{blk}.dw.w"].T).reshape(dim, 1, 7, 7)
            sd[lp + ".dwconv.bias"] = t[f"blk{blk}.dw.b"].copy()
            sd[lp + ".layernorm.weight"], sd[lp + ".layernorm.bias"] = t[f"blk{blk}.ln.w"].copy(), t[f"blk{blk}.ln.b"].copy()
            sd[lp + ".pwconv1.weight"], sd[lp + ".pwconv1.bias"] = _unpack_linear(t, f"blk{blk}.pw1", dim)
            sd[lp + ".pwconv2.weight"], sd[lp + ".pwconv2.bias"] = _unpack_linear(t, f"blk{blk}.pw2", 4 * dim)
            sd[lp + ".layer_scale_parameter"] = np.ones(dim, np.float32)
            blk += 1
    v = "vitstr.vit"
    w, b = _unpack_linear(t, "vit.proj", 512)
    sd[v + ".embeddings.patch_embeddings.projection.weight"], sd[v + ".embeddings.patch_embeddings.projection.bias"] = w.reshape(192, 512, 1, 1), b
    sd[v + ".embeddings.position_embeddings"] = np.concatenate([np.zeros((1, 192), np.float32), t["vit.pos"]], 0)[None]
    L = 0
    while f"vit{L}.qkv.w" in t:
        lp = f"{v}.encoder.layer.{L}"
        w, b = _unpack_linear(t, f"vit{L}.qkv", 192)
        for k, name in enumerate(("query", "key", "value")):
            scale = np.float32(8.0) if k == 0 else np.float32(1.0)
            sd[f"{lp}.attention.attention.{name}.weight"] = w[k * 192:(k + 1) * 192] * scale
            sd[f"{lp}.attention.attention.{name}.bias"] = b[k * 192:(k + 1) * 192] * scale
        sd[lp + ".attention.output.dense.weight"], sd[lp + ".attention.output.dense.bias"] = _unpack_linear(t, f"vit{L}.proj", 192)
        sd[lp + ".intermediate.dense.weight"], sd[lp + ".intermediate.dense.bias"] = _unpack_linear(t, f"vit{L}.fc1", 192)
        sd[lp + ".output.dense.weight"], sd[lp + ".output.dense.bias"] = _unpack_linear(t, f"vit{L}.fc2", 768)
        sd[lp + ".layernorm_before.weight"], sd[lp + ".layernorm_before.bias"] = t[f"vit{L}.ln1.w"].copy(), t[f"vit{L}.ln1.b"].copy()
        sd[lp + ".layernorm_after.weight"], sd[lp + ".layernorm_after.bias"] = t[f"vit{L}.ln2.w"].copy(), t[f"vit{L}.ln2.b"].copy()
        L += 1
    sd[v + ".layernorm.weight"], sd[v + ".layernorm.bias"] = t["vit.ln.w"].copy(), t["vit.ln.b"].copy()
    sd["vitstr.classifier.weight"], sd["vitstr.classifier.bias"] = _unpack_linear(t, "cls", 192)
    return sd


@torch.no_grad()
def lore_resnet18_from_blob(t: Dict[str, np.ndarray], x: torch.Tensor) -> Dict[str, torch.Tensor]:
    """t = read_blob(pack_lore_resnet18(sd)); x fp32 [N,3,H,W] (H, W multiples of 64) -> {'maps': the packed 24-wide head map
    [N,24,H/4,W/4] (hm NOT sigmoid-ed here), 'ax' / 'cr': dense [N,256,H/4,W/4]} computed FROM THE PACKED TENSORS in the order
    csrc/lore_net.cu build_r18 plans them: the transposed convs as 3x3 convs to 4 x 256 pixel-shuffled channels, the six first
    head convs as one conv, the 64 -> 64 convs on channel slices [hm | reg | wh | st | ax | cr], the small heads' last 1x1 as one
    block-diagonal conv over the first 256 hidden channels (reg's slice is its first conv's output)."""
    def conv(name, inp, cin, k, stride=1, relu=False):
        w, b = _unpack_conv(t, name, cin, k)
        y = F.conv2d(inp, w, b, stride=stride, padding=k // 2)
        return F.relu(y) if relu else y

    ws = torch.from_numpy(t["stem.w"].astype(np.float32)).reshape(64, 7, 8, 4)[:, :, :7, :3].permute(0, 3, 1, 2).contiguous()
    y = F.relu(F.conv2d(x.float(), ws, torch.from_numpy(t["stem.b"][:64].copy()), stride=2, padding=3))
    xs = [F.max_pool2d(y, 3, 2, 1)]
    y, cin = xs[0], 64
    for L, cout in zip(range(1, 5), (64, 128, 256, 256)):
        for B in range(2):
            stride = 2 if B == 0 else 1
            p = f"layer{L}.{B}"
            o = conv(p + ".conv1", y, cin, 3, stride, relu=True)
            o = conv(p + ".conv2", o, cout, 3)
            if (p + ".down.w") in t:
                y = conv(p + ".down", y, cin, 1, stride)
            y = F.relu(o + y)
            cin = cout
        xs.append(y)
    top = xs[4]
    for i, (name, skip) in enumerate((("adaption3", xs[3]), ("adaption2", xs[2]), ("adaption1", xs[1]), ("adaption0", xs[0]))):
        up = conv(f"up{i + 1}", top, 256, 3, relu=True)  # [N, 4*256, h, w], row (py*2 + px)*256 + co
        n, _, h, w = up.shape
        up = up.reshape(n, 2, 2, 256, h, w).permute(0, 3, 4, 1, 5, 2).reshape(n, 256, 2 * h, 2 * w)
        top = conv(name, skip, skip.shape[1], 1) + up
    feat = conv("adaptionU1", top, 256, 1)
    a = conv("heads.conv1", feat, 256, 3, relu=True)  # 6 x 64
    hid = []
    for h, head in enumerate(("hm", "reg", "wh", "st", "ax", "cr")):
        v = a[:, 64 * h: 64 * (h + 1)]
        if head != "reg":
            for j in (2, 4, 6):
                v = conv(f"heads.{head}.{j}", v, 64, 3, relu=True)
        hid.append(v)
    return {"maps": conv("heads.out", torch.cat(hid[:4], 1), 256, 1), "ax": conv("ax.out", hid[4], 64, 1), "cr": conv("cr.out", hid[5], 64, 1)}
