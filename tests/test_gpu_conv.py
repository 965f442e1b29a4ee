"""conv_igemm_tcgen05 (operator level) vs torch fp32 conv2d on the same fp16-rounded operands."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from pdf_table_b200 import weights

pytestmark = pytest.mark.gpu

CASES = [
    # name, N, H, W, Cin, Cout, k, stride, pad, residual, act
    ("1x1_flat_partial_tile", 2, 9, 13, 64, 64, 1, 1, 0, False, 0),
    ("1x1_flat_k96_gelu", 1, 8, 75, 96, 384, 1, 1, 0, False, 2),
    ("1x1_flat_k384_res", 1, 8, 75, 384, 96, 1, 1, 0, True, 0),
    ("3x3_s1_res_relu", 2, 20, 24, 64, 128, 3, 1, 1, True, 1),
    ("3x3_s1_k2304", 1, 30, 30, 256, 64, 3, 1, 1, False, 0),
    ("3x3_s2", 2, 32, 48, 64, 128, 3, 2, 1, False, 1),
    ("1x1_s2", 2, 32, 48, 64, 128, 1, 2, 0, False, 0),
    ("3x3_ntiles2", 1, 16, 16, 128, 512, 3, 1, 1, False, 1),
    ("3x3_cout40", 1, 12, 20, 32, 40, 3, 1, 1, False, 0),
    ("3x3_many_tiles", 4, 120, 120, 64, 64, 3, 1, 1, True, 1),
    ("3x3_c16", 1, 24, 40, 16, 32, 3, 1, 1, False, 1),
    ("3x3_c96_cout32", 2, 40, 56, 96, 32, 3, 1, 1, False, 1),     # PP-OCRv4 det head / neck shape (BK 32, three channel blocks)
    ("3x3_c64_cout32_odd", 3, 37, 29, 64, 32, 3, 1, 1, False, 0),  # the offset / mask conv of a deformable layer, ragged tiles
]


@pytest.mark.parametrize("case", CASES, ids=[c[0] for c in CASES])
def test_conv_parity(post_engine, case):
    name, N, H, W, Cin, Cout, k, stride, pad, use_res, act = case
    rng = np.random.default_rng(abs(hash(name)) % (2 ** 31))
    x = torch.from_numpy(rng.standard_normal((N, Cin, H, W)).astype(np.float32)).half().float()
    w = torch.from_numpy((rng.standard_normal((Cout, Cin, k, k)) / np.sqrt(Cin * k * k)).astype(np.float32)).half().float()
    b = torch.from_numpy(rng.standard_normal(Cout).astype(np.float32))
    ref = F.conv2d(x, w, b, stride=stride, padding=pad)
    res = None
    if use_res:
        res = torch.from_numpy(rng.standard_normal(tuple(ref.shape)).astype(np.float32)).half().float()
        ref = ref + res
    if act == 1:
        ref = F.relu(ref)
    elif act == 2:
        ref = F.gelu(ref)
    wp, bp = weights.pack_conv(w.numpy(), b.numpy())
    eng = post_engine
    x_nhwc = eng.nchw_to_nhwc_f16(x.cuda())
    res_nhwc = eng.nchw_to_nhwc_f16(res.cuda()) if use_res else None
    out = eng.conv2d_nhwc(x_nhwc, torch.from_numpy(wp).cuda(), torch.from_numpy(bp).cuda(), Cout, k, stride, pad,
                          residual=res_nhwc, act=act)
    got = eng.nhwc_f16_to_nchw(out).cpu()
    eng.sync()
    err = (got - ref).abs()
    tol = 2e-3 * ref.abs().clamp(min=1.0)  # fp16 output rounding (2^-11 relative) + accumulation order
    assert bool((err <= tol).all()), f"{name}: max err {float(err.max())} at {int(err.argmax())}"
