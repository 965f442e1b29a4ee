"""dv_lore_decode / dv_lore_gather_logi (CUDA) vs the reference-generated golden rows and the oracle restatement."""
import os

import numpy as np
import pytest
import torch

from oracle import lore_decode_ref
from pdf_table_b200 import synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DECODE_CASES = [("t0", 0, 128, 128), ("t1", 1, 128, 128), ("t2", 2, 96, 160), ("t3", 3, 256, 256)]


def _cuda(m, keys):
    return [torch.from_numpy(m[k])[None].cuda() for k in keys]


def _inv_affine(meta):
    return lore_decode_ref.affine_matrix([np.float32(meta[0]), np.float32(meta[1])], np.float32(meta[2]), int(meta[6]), int(meta[5]), True)


def test_lore_decode_reference_golden(post_engine):
    g = np.load(os.path.join(GOLDEN, "lore_decode.npz"))
    for name, idx, h, w in DECODE_CASES:
        m = synth.lore_planted_maps(idx, h, w)
        meta = g[name + "_meta"]
        hm, reg, wh, st, ax, cr = _cuda(m, ("hm", "reg", "wh", "st", "ax", "cr"))
        dec = post_engine.lore_decode(hm, reg, wh, st, _inv_affine(meta)[None], check_overflow=True)
        logi = post_engine.lore_gather_logi(ax, cr, dec)
        post_engine.sync()
        assert dec["overflow"] == 0
        n = int(dec["counts"].cpu()[0])
        res = g[name + "_results"]
        assert n == len(g[name + "_logi_feat"]), name
        np.testing.assert_array_equal(dec["polygons"].cpu().numpy()[0, :n], res[:n, :8], err_msg=name)
        np.testing.assert_array_equal(dec["scores"].cpu().numpy()[0, :n], res[:n, 8], err_msg=name)
        np.testing.assert_array_equal(dec["dets_feat"].cpu().numpy()[0, :n], g[name + "_dets_feat"], err_msg=name)
        np.testing.assert_array_equal(logi.cpu().numpy()[0, :n], g[name + "_logi_feat"], err_msg=name)
        # rows below vis_thresh: the gated cells (>= 0.2 before the x0.4 penalty) keep their relative reference order
        rows = int(dec["rows"].cpu()[0])
        assert rows >= n
        tail = dec["scores"].cpu().numpy()[0, n:rows]
        assert (tail < 0.2).all() and (np.diff(tail) <= 0).all()


def test_lore_decode_batch_packed_layout_vs_oracle(post_engine):
    """Three images in one call through the packed NHWC x24 layout the network writes; per-image affine."""
    cases = [(10, (700, 900)), (11, (1024, 1024)), (12, (400, 1300))]
    h = w = 128
    packed = np.zeros((len(cases), h, w, 24), np.float32)
    want, trans = [], []
    for i, (idx, (sh, sw)) in enumerate(cases):
        m = synth.lore_planted_maps(idx, h, w, with_feat=False)
        packed[i, :, :, 0:2] = m["hm"].transpose(1, 2, 0)
        packed[i, :, :, 2:4] = m["reg"].transpose(1, 2, 0)
        packed[i, :, :, 4:12] = m["wh"].transpose(1, 2, 0)
        packed[i, :, :, 12:20] = m["st"].transpose(1, 2, 0)
        meta = np.array([int(sw / 2.0), int(sh / 2.0), max(sh, sw), 4 * h, 4 * w, h, w])
        z = np.zeros((1, h, w), np.float32)
        want.append(lore_decode_ref.lore_decode(m["hm"], m["reg"], m["wh"], m["st"], z, z, meta))
        trans.append(want[-1]["trans"])
    dec = post_engine.lore_decode(torch.from_numpy(packed).cuda(), None, None, None, np.stack(trans))
    post_engine.sync()
    for i in range(len(cases)):
        n = int(dec["counts"].cpu()[i])
        assert n == len(want[i]["polygons"]) and n > 0
        np.testing.assert_array_equal(dec["polygons"].cpu().numpy()[i, :n], want[i]["polygons"])
        np.testing.assert_array_equal(dec["dets_feat"].cpu().numpy()[i, :n], want[i]["dets_feat"])
        rows = int(dec["rows"].cpu()[i])
        np.testing.assert_array_equal(dec["cr_idx"].cpu().numpy()[i, :rows], want[i]["cc_match"][:rows])
        np.testing.assert_array_equal(dec["ax_idx"].cpu().numpy()[i, :rows], want[i]["cell_inds"][want[i]["order"]][:rows])


def test_lore_decode_edge_cases(post_engine):
    h = w = 80
    z2, z8 = torch.zeros(1, 2, h, w).cuda(), torch.zeros(1, 8, h, w).cuda()
    eye = np.array([[[1.0, 0, 0], [0, 1.0, 0]]])
    dec = post_engine.lore_decode(z2, z2, z8, z8, eye)  # empty map: nothing above the gates
    post_engine.sync()
    assert int(dec["counts"].cpu()[0]) == 0 and int(dec["rows"].cpu()[0]) == 0
    # without wiz_rev (wireless / ptn configurations): top-K order is kept, no penalty, cc_match from the raw corners
    m = synth.lore_planted_maps(5, 96, 96, with_feat=False)
    hm, reg, wh, st = _cuda(m, ("hm", "reg", "wh", "st"))
    meta = np.array([200, 150, 400, 384, 384, 96, 96])
    z = np.zeros((1, 96, 96), np.float32)
    want = lore_decode_ref.lore_decode(m["hm"], m["reg"], m["wh"], m["st"], z, z, meta, wiz_rev=False, vis_thresh=0.35)
    dec = post_engine.lore_decode(hm, reg, wh, st, want["trans"][None], wiz_rev=False, vis_thresh=0.35)
    post_engine.sync()
    n = int(dec["counts"].cpu()[0])
    assert n == len(want["polygons"]) and n > 0
    np.testing.assert_array_equal(dec["polygons"].cpu().numpy()[0, :n], want["polygons"])
    np.testing.assert_array_equal(dec["cr_idx"].cpu().numpy()[0, :n], want["cc_match"][:n])
    # K smaller than the number of gated peaks: the best K survive
    want = lore_decode_ref.lore_decode(m["hm"], m["reg"], m["wh"], m["st"], z, z, meta, K=7, MK=9)
    dec = post_engine.lore_decode(hm, reg, wh, st, want["trans"][None], K=7, MK=9)
    post_engine.sync()
    n = int(dec["counts"].cpu()[0])
    assert n == len(want["polygons"])
    np.testing.assert_array_equal(dec["polygons"].cpu().numpy()[0, :n], want["polygons"])
    np.testing.assert_array_equal(dec["scores"].cpu().numpy()[0, :n], want["scores"])
