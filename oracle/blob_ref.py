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
