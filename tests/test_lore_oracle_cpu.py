"""Lore oracle restatements vs the golden fixtures produced by the reference's own modules
(oracle/gen_golden_lore.py, run in the build container)."""
import os

import numpy as np
import torch

from oracle import lore_decode_ref, lore_net_ref, lore_processor_ref
from pdf_table_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DECODE_CASES = [("t0", 0, 128, 128), ("t1", 1, 128, 128), ("t2", 2, 96, 160), ("t3", 3, 256, 256)]


def test_lore_network_oracle_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "lore_dla34_seed0.npz"))
    sd = synth.lore_dla34_state_dict(0)
    for tv in (True, False):  # torchvision.ops.deform_conv2d and its plain-torch restatement
        out = lore_net_ref.lore_dla34_forward(sd, torch.from_numpy(g["x"]), use_torchvision=tv)
        for k in ("hm", "st", "wh", "ax", "cr", "reg"):
            np.testing.assert_allclose(out[k].numpy(), g[k], atol=2e-5, rtol=0, err_msg=f"{k} tv={tv}")


def test_deform_conv_restatement_matches_torchvision():
    from torchvision.ops import deform_conv2d

    gen = torch.Generator().manual_seed(0)
    x = torch.randn(2, 16, 12, 20, generator=gen)
    off = torch.randn(2, 18, 12, 20, generator=gen) * 3  # samples far outside the image included
    mask = torch.rand(2, 9, 12, 20, generator=gen)
    w, b = torch.randn(8, 16, 3, 3, generator=gen), torch.randn(8, generator=gen)
    want = deform_conv2d(x, off, w, b, padding=(1, 1), mask=mask)
    got = lore_net_ref.deform_conv2d_ref(x, off, w, b, mask)
    np.testing.assert_allclose(got.numpy(), want.numpy(), atol=1e-5, rtol=0)


def test_lore_decode_oracle_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "lore_decode.npz"))
    for name, idx, h, w in DECODE_CASES:
        m = synth.lore_planted_maps(idx, h, w)
        out = lore_decode_ref.lore_decode(m["hm"], m["reg"], m["wh"], m["st"], m["ax"], m["cr"], g[name + "_meta"])
        res = g[name + "_results"]
        n = len(g[name + "_logi_feat"])
        assert n > 0 and len(out["polygons"]) == n, name
        np.testing.assert_array_equal(out["polygons"], res[:n, :8], err_msg=name)
        np.testing.assert_array_equal(out["scores"], res[:n, 8], err_msg=name)
        np.testing.assert_array_equal(out["dets_feat"], g[name + "_dets_feat"], err_msg=name)
        np.testing.assert_allclose(out["logi_feat"], g[name + "_logi_feat"], atol=1e-6, rtol=0, err_msg=name)
        # every row the x0.4 penalty could have produced from a >= 0.2 cell is also in reference order
        m08 = int((res[:, 8] >= 0.08).sum())
        np.testing.assert_array_equal(out["all_scores"][:m08], res[:m08, 8], err_msg=name)


def test_lore_decode_known_answer():
    """One cell whose four corners each have a detected corner point inside it: the corners snap to the corner
    points, the score is kept, and the polygon maps back to source pixels through the inverse affine."""
    h = w = 80
    hm = np.zeros((2, h, w), np.float32)
    reg = np.full((2, h, w), 0.5, np.float32)
    wh = np.zeros((8, h, w), np.float32)
    st = np.zeros((8, h, w), np.float32)
    hm[0, 40, 40] = 0.9
    corners = [(30, 30), (50, 30), (50, 50), (30, 50)]
    for k, (x, y) in enumerate(corners):
        hm[1, y, x] = 0.8 - 0.01 * k
        wh[2 * k, 40, 40] = 40.5 - (x + 0.5) + (1.0 if x < 40 else -1.0) * -1.5  # predicted corner 1.5 px outside
        wh[2 * k + 1, 40, 40] = 40.5 - (y + 0.5) + (1.0 if y < 40 else -1.0) * -1.5
        # the corner's own 4-point box reaches 3 px diagonally into each neighbouring cell
        for q, (dx, dy) in enumerate([(-3, -3), (3, -3), (3, 3), (-3, 3)]):
            st[2 * q, y, x], st[2 * q + 1, y, x] = -dx, -dy
    ax = np.zeros((4, h, w), np.float32)
    cr = np.zeros((4, h, w), np.float32)
    meta = np.array([160, 160, 320, 320, 320, 80, 80])
    out = lore_decode_ref.lore_decode(hm, reg, wh, st, ax, cr, meta)
    assert len(out["polygons"]) == 1
    np.testing.assert_array_equal(out["scores"], np.float32([0.9]))
    np.testing.assert_array_equal(out["dets_feat"][0], [30, 30, 50, 30, 50, 50, 30, 50])
    np.testing.assert_allclose(out["polygons"][0], np.float32([30.5, 30.5, 50.5, 30.5, 50.5, 50.5, 30.5, 50.5]) * 4, atol=1e-4)


def test_lore_processor_oracle_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "lore_processor_seed0.npz"))
    sd = synth.lore_processor_state_dict(0)
    for n in (1, 7, 64, 200):
        logic, stacked = lore_processor_ref.lore_processor_forward(sd, torch.from_numpy(g[f"n{n}_feat"]))
        np.testing.assert_allclose(logic.numpy(), g[f"n{n}_logic"], atol=1e-5, rtol=0)
        np.testing.assert_allclose(stacked.numpy(), g[f"n{n}_stacked"], atol=1e-5, rtol=0)


def test_round_logic_is_half_down():
    x = np.float32([0.5, 1.5, 2.5000002, 3.49, 0.0, 7.51])
    np.testing.assert_array_equal(lore_decode_ref.round_logic(x), np.float32([0, 1, 3, 3, 0, 8]))


def test_picodet_oracle_matches_reference_golden():
    from oracle import picodet_ref

    g = np.load(os.path.join(GOLDEN, "picodet_post.npz"))
    cases = [("en", 0, 5, (1100, 850), 12), ("ch", 1, 10, (1600, 1200), 20), ("table", 2, 1, (700, 1000), 4),
             ("empty", 3, 5, (800, 608), 0), ("dense", 4, 5, (2000, 1500), 60)]
    for name, idx, c, (oh, ow), nobj in cases:
        s, b = synth.picodet_planted_outputs(idx, c, n_objects=nobj)
        got = picodet_ref.picodet_decode(s, b, [oh, ow], [800.0 / oh, 608.0 / ow], [800, 608])[0]
        np.testing.assert_array_equal(got, g[name], err_msg=name)
    # SURVEY.md 8c known-answer case: one anchor (level 1, index 500) scoring 0.9 on class 3, DFL one-hot at bin 5
    scores = [np.zeros((1, hw, 5), np.float32) for hw in (7600, 1900, 475, 130)]
    boxes = [np.zeros((1, hw, 32), np.float32) for hw in (7600, 1900, 475, 130)]
    scores[1][0, 500, 3] = 0.9
    boxes[1][0, 500].reshape(4, 8)[:, 5] = 100.0
    got = picodet_ref.picodet_decode(scores, boxes, [1600, 1216], [0.5, 0.5], [800, 608])[0]
    assert got.shape == (1, 6) and got[0, 0] == 3
    np.testing.assert_allclose(got[0, 2:], [48, 272, 368, 592], atol=1e-3)


def test_picodet_network_oracle_matches_reference_golden():
    from oracle import picodet_net_ref

    g = np.load(os.path.join(GOLDEN, "picodet_net_seed0.npz"))
    bb, nk, hd = synth.picodet_state_dicts(0, 5)
    s, d = picodet_net_ref.picodet_forward(bb, nk, hd, torch.from_numpy(g["x"]), 5)
    for lvl in range(4):
        np.testing.assert_allclose(s[lvl].numpy(), g[f"scores{lvl}"], atol=1e-6, rtol=0)
        np.testing.assert_allclose(d[lvl].numpy(), g[f"dfl{lvl}"], atol=1e-5, rtol=0)


def test_picodet_graph_program_is_consistent():
    """The lowered program: every op reads tensors / slices that an earlier op wrote, slices stay inside their buffers."""
    from pdf_table_b200 import picodet_graph as G

    bb, nk, hd = synth.picodet_state_dicts(0, 5)
    blob, meta = G.build_picodet(bb, nk, hd, 5)
    tensors, ops = blob["graph.tensors"], blob["graph.ops"]
    written = {0: [(0, 3)]}
    for code, in_t, in_coff, in_c, out_t, out_coff, out_c, k, stride, act, w, aux in ops:
        assert in_coff + in_c <= tensors[in_t][0]
        assert any(a <= in_coff and in_coff + in_c <= b for a, b in _merge(written.get(in_t, []))), (code, in_t, in_coff, in_c)
        if code == G.OP_ADD:
            assert aux in written
        if code != G.OP_HEAD:
            assert out_coff + out_c <= tensors[out_t][0]
            written.setdefault(out_t, []).append((out_coff, out_coff + out_c))
    assert len([o for o in ops if o[0] == G.OP_HEAD]) == 4
    assert set(meta["features"]) == {"c3", "c4", "c5", "p3", "p4", "p5", "p6"}


def _merge(iv):
    out = []
    for a, b in sorted(iv):
        if out and a <= out[-1][1]:
            out[-1] = (out[-1][0], max(out[-1][1], b))
        else:
            out.append((a, b))
    return out


def test_centernet_oracle_matches_reference_golden():
    from oracle import centernet_ref

    g = np.load(os.path.join(GOLDEN, "centernet_dla34_seed0.npz"))
    out = centernet_ref.centernet_dla34_forward(synth.centernet_dla34_state_dict(0), torch.from_numpy(g["x"]))
    for k in ("hm", "v2c", "c2v", "reg"):
        np.testing.assert_allclose(out[k].numpy(), g[k], atol=2e-5 * max(1.0, float(np.abs(g[k]).max())), rtol=0, err_msg=k)
    d = np.load(os.path.join(GOLDEN, "centernet_decode.npz"))
    for name, idx, h, w, (sh, sw) in [("c0", 20, 128, 128, (600, 800)), ("c1", 21, 128, 160, (1024, 1280)), ("c2", 22, 256, 256, (1500, 1100))]:
        m = synth.lore_planted_maps(idx, h, w, with_feat=False)
        got = centernet_ref.centernet_decode(m["hm"], m["reg"], m["wh"], m["st"], [sw / 2.0, sh / 2.0], max(sh, sw) * 1.0, h, w)
        np.testing.assert_array_equal(got, d[name], err_msg=name)


def test_point_in_polygon_restatement_agrees_with_cv2():
    """The strict point-in-polygon test that stands in for shapely's Point.within(Polygon) (absent from the image, parity
    unpinned) cross-checked against an independent implementation, cv2.pointPolygonTest (+1 inside, 0 on the boundary, -1
    outside): random convex and concave quads, random points, and points planted exactly on vertices and (integer quads) edges."""
    import cv2

    rng = np.random.default_rng(21)
    checked = boundary = 0
    for trial in range(300):
        ang = np.sort(rng.uniform(0, 2 * np.pi, 4))
        rad = rng.uniform(20, 100, 4) if trial % 3 else np.array([90.0, 25.0, 90.0, 90.0])  # every third quad is concave
        quad = (np.stack([np.cos(ang), np.sin(ang)], 1) * rad[:, None] + 200).astype(np.float32)
        if trial % 2:
            quad = np.rint(quad).astype(np.float32)  # integer corners: exact boundary hits exist
        pts = [tuple(rng.uniform(80, 320, 2)) for _ in range(30)]
        pts += [tuple(map(float, quad[k])) for k in range(4)]                              # vertices
        if trial % 2:  # edge mid-points: exactly on the edge only for integer corners (cv2 forms its differences in float32)
            pts += [tuple(map(float, (quad[k].astype(np.float64) + quad[(k + 1) % 4]) / 2)) for k in range(4)]
        for px, py in pts:
            want = cv2.pointPolygonTest(quad.reshape(-1, 1, 2), (float(px), float(py)), False)
            if want != 0 and abs(cv2.pointPolygonTest(quad.reshape(-1, 1, 2), (float(px), float(py)), True)) < 1e-3:
                continue  # closer to the boundary than cv2's float32 arithmetic resolves
            got = lore_decode_ref.point_strictly_in_polygon(float(px), float(py), quad)
            if want == 0:
                boundary += 1
            assert got == (want > 0), (trial, px, py, want)
            checked += 1
    assert checked > 10000 and boundary > 500


def test_polygon_area_and_length_stand_in_agrees_with_cv2():
    """shapely's Polygon.area / .length stand-in (DB unclip distance, parity unpinned) against cv2.contourArea / cv2.arcLength."""
    import cv2

    from oracle import db_post_ref

    rng = np.random.default_rng(22)
    for trial in range(200):
        ang = np.sort(rng.uniform(0, 2 * np.pi, 4))
        quad = np.rint(np.stack([np.cos(ang), np.sin(ang)], 1) * rng.uniform(5, 300, 4)[:, None] + 400).astype(np.int32)
        poly = db_post_ref.ShapelyPolygonStandIn(quad)
        assert poly.area == cv2.contourArea(quad.reshape(-1, 1, 2))  # integer shoelace sums are exact in both
        assert abs(poly.length - cv2.arcLength(quad.reshape(-1, 1, 2).astype(np.float32), True)) <= 1e-4 * max(poly.length, 1.0)


def test_picodet_lowering_computes_the_reference_function():
    """The graph program + packed weights that the engine executes (picodet_graph.build_picodet), run op by op on CPU with the
    executor's semantics (oracle/graph_interp.py), reproduce the oracle restatement of LCNet + CSP-PAN + PicoHead: features of
    every stage and the eight head outputs.  Checks BatchNorm folding, packing, slice-concatenations and op order on CPU."""
    from oracle import graph_interp, picodet_net_ref
    from pdf_table_b200 import picodet_graph as G

    bb, nk, hd = synth.picodet_state_dicts(0, 5)
    blob, meta = G.build_picodet(bb, nk, hd, 5)
    x = torch.from_numpy(np.random.default_rng(31).standard_normal((2, 3, 320, 256)).astype(np.float32))
    ws, wd, feats = picodet_net_ref.picodet_forward(bb, nk, hd, x, 5, return_features=True)
    for fp16_act, tol in ((False, 2e-4), (True, 6e-4)):  # fp16 weights only (measured 7e-5) / + fp16 activation buffers as on the device (2.4e-4)
        tens, heads = graph_interp.run_program(blob, x, fp16_activations=fp16_act)
        for name, tid in meta["features"].items():
            want = feats[name]
            rel = float((tens[tid] - want).abs().max()) / max(1.0, float(want.abs().max()))
            assert rel < tol, (fp16_act, name, rel)
        assert sorted(heads) == [0, 1, 2, 3]
        for lvl in range(4):
            s, d = heads[lvl]
            assert s.shape == ws[lvl].shape and d.shape == wd[lvl].shape
            assert float((s - ws[lvl]).abs().max()) < tol
            assert float((d - wd[lvl]).abs().max()) < tol * max(1.0, float(wd[lvl].abs().max()))


def test_pp_rec_graph_program_reproduces_the_oracle_on_cpu():
    """The lowering of the PP-OCRv4 recogniser (pp_rec_graph.py: rep-layer affines folded / kept as post-activation affines,
    BatchNorm folding, (1,3) convs as unfold + GEMM, channel padding 60 -> 64, attention scale folded into q, concatenation
    slices, CTC head padding) run op by op with the executor's semantics equals the oracle of the published architecture up to
    the fp16 rounding of the packed weights (and of the activation buffers)."""
    import torch

    from oracle import graph_interp, pp_rec_ref
    from pdf_table_b200 import pp_rec_graph, synth

    sd = synth.pp_ocrv4_rec_state_dict(0, 97)
    blob, meta = pp_rec_graph.build_pp_rec(sd)
    assert meta["n_class"] == 97 and int(blob["graph.meta"][5]) == 1
    for w in (320, 325, 96):
        x = torch.from_numpy(np.random.default_rng(1).standard_normal((2, 3, 48, w)).astype(np.float32))
        want, want_logits = pp_rec_ref.pp_rec_forward(sd, x, return_logits=True)
        for fp16_act, tol in ((False, 4e-2), (True, 6e-2)):
            _, heads = graph_interp.run_program(blob, x, fp16_activations=fp16_act)
            assert tuple(heads["probs"].shape) == tuple(want.shape)
            assert float((heads["logits"] - want_logits).abs().max()) < tol  # logit std ~4.7
            assert float((heads["probs"] - want).abs().max()) < 1e-2


def test_pp_det_graph_program_reproduces_the_oracle_on_cpu():
    """The lowering of the PP-OCRv4 detector (pp_det_graph.py: stride-2 rep layers without activation, tap conv folded into the
    neck's ins_conv, RSE shortcut SE with the slope-0.2 hardsigmoid folded into conv2, top-down sums, up-sampling into the
    concatenation slices, padded head channels, transposed convs as GEMM + pixel shuffle) run op by op with the executor's
    semantics equals the oracle of the published architecture up to the fp16 rounding of the packed weights / buffers."""
    import torch

    from oracle import graph_interp, pp_det_ref
    from pdf_table_b200 import pp_det_graph, synth

    sd = synth.pp_ocrv4_det_state_dict(0)
    blob, meta = pp_det_graph.build_pp_det(sd)
    assert int(blob["graph.meta"][5]) == 3
    for h, w in ((96, 160), (64, 64)):
        x = torch.from_numpy(np.random.default_rng(2).standard_normal((2, 3, h, w)).astype(np.float32))
        want, fuse = pp_det_ref.pp_det_forward(sd, x, return_fuse=True)
        for fp16_act, tol in ((False, 5e-3), (True, 8e-3)):
            tens, heads = graph_interp.run_program(blob, x, fp16_activations=fp16_act)
            assert tuple(heads["prob"].shape) == tuple(want.shape) == (2, 1, h, w)
            assert float((tens[meta["fuse"]] - fuse).abs().max()) < 1e-2 * float(fuse.abs().max())
            assert float((heads["prob"] - want).abs().max()) < tol


def test_lore_wireless_network_oracle_matches_reference_golden():
    """oracle/lore_wireless_ref.py vs LoreDetectModel (lore/lore_detector.py) run by oracle/gen_golden_lore_wireless.py."""
    from oracle import lore_wireless_ref

    g = np.load(os.path.join(GOLDEN, "lore_resnet18_seed0.npz"))
    out = lore_wireless_ref.lore_resnet18_forward(synth.lore_resnet18_state_dict(0), torch.from_numpy(g["x"]))
    for k in ("hm", "st", "wh", "ax", "cr", "reg"):
        np.testing.assert_allclose(out[k].numpy(), g[k], atol=2e-5, rtol=0, err_msg=k)


def test_lore_wireless_decode_oracle_matches_reference_golden():
    """process_detect_output(upper_left=True, wiz_rev=False): the decode in the wireless configuration's frame."""
    g = np.load(os.path.join(GOLDEN, "lore_resnet18_seed0.npz"))
    m = synth.lore_planted_maps(4, 192, 192)
    out = lore_decode_ref.lore_decode(m["hm"], m["reg"], m["wh"], m["st"], m["ax"], m["cr"], g["dec_meta"], upper_left=True, wiz_rev=False)
    res = g["dec_results"]
    n = len(g["dec_logi_feat"])
    assert n > 0 and len(out["polygons"]) == n
    np.testing.assert_array_equal(out["polygons"], res[:n, :8])
    np.testing.assert_array_equal(out["scores"], res[:n, 8])
    np.testing.assert_array_equal(out["dets_feat"], g["dec_dets_feat"])
    np.testing.assert_allclose(out["logi_feat"], g["dec_logi_feat"], atol=1e-6, rtol=0)


def test_packed_lore_resnet18_blob_computes_the_reference_function():
    """pack_lore_resnet18 on CPU: BatchNorm folding, the ConvTranspose 4x4 s2 -> 3x3 conv + pixel shuffle identity, the fused first
    head conv, the block-diagonal last 1x1 -- the network run from the packed tensors equals the oracle up to fp16 weight rounding."""
    from oracle import blob_ref, lore_wireless_ref
    from pdf_table_b200 import weights

    sd = synth.lore_resnet18_state_dict(0)
    w3, _ = weights.deconv4x4_as_conv3x3(sd["deconv_layers1.0.weight"])
    xin = torch.from_numpy(np.random.default_rng(2).standard_normal((1, 256, 5, 7)).astype(np.float32))
    y = torch.nn.functional.conv2d(xin, torch.from_numpy(w3), padding=1).reshape(1, 2, 2, 256, 5, 7).permute(0, 3, 4, 1, 5, 2).reshape(1, 256, 10, 14)
    ref = torch.nn.functional.conv_transpose2d(xin, torch.from_numpy(sd["deconv_layers1.0.weight"]), stride=2, padding=1)
    assert float((y - ref).abs().max()) < 1e-4  # exact up to the summation order
    t = blob_ref.read_blob(weights.pack_lore_resnet18(sd))
    assert t["up1.w"].shape == (1024, 9 * 256) and t["heads.conv1.w"].shape == (384, 9 * 256) and t["heads.out.w"].shape == (24, 256)
    g = np.load(os.path.join(GOLDEN, "lore_resnet18_seed0.npz"))
    x = torch.from_numpy(g["x"])
    got = blob_ref.lore_resnet18_from_blob(t, x)
    worst = 0.0
    for k, sl in (("hm", slice(0, 2)), ("reg", slice(2, 4)), ("wh", slice(4, 12)), ("st", slice(12, 20))):
        worst = max(worst, float(np.abs(got["maps"][:, sl].numpy() - g[k]).max()) / max(1.0, float(np.abs(g[k]).max())))
    for k in ("ax", "cr"):
        worst = max(worst, float(np.abs(got[k].numpy() - g[k]).max()) / max(1.0, float(np.abs(g[k]).max())))
    assert worst < 3e-3, worst  # fp16 rounding of the packed weights
    assert float(got["maps"][:, 20:].abs().max()) == 0.0  # the four pad columns
