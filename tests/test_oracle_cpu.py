"""Oracle restatements vs the golden fixtures produced by the reference's own modules
(oracle/gen_golden.py, run in the build container)."""
import os

import numpy as np
import pytest
import torch

from oracle import ctc_ref, dbnet_ref
from pdf_table_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def test_ctc_oracle_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "ctc_decode.npz"))
    character = list(g["character"])
    for case in ("known", "rand_T40", "rand_T7", "rand_T160", "rand_T300", "edge"):
        res = ctc_ref.ctc_decode_text(g[f"{case}.preds"], character)
        assert [r[0] for r in res] == list(g[f"{case}.text"]), case
        np.testing.assert_array_equal(np.array([r[1] for r in res]), g[f"{case}.conf"], err_msg=case)
    # SURVEY.md 8c known-answer vector
    assert ctc_ref.ctc_decode_text(g["known.preds"], character)[0][0] == "012 "


def test_dbnet_oracle_matches_reference_golden():
    g = np.load(os.path.join(GOLDEN, "dbnet_r18_seed0.npz"))
    sd = synth.dbnet_r18_state_dict(0)
    prob = dbnet_ref.dbnet_r18_forward(sd, torch.from_numpy(g["x"])).numpy()
    np.testing.assert_allclose(prob, g["prob"], atol=1e-6, rtol=0)


def test_synth_weights_deterministic():
    a, b = synth.dbnet_r18_state_dict(0), synth.dbnet_r18_state_dict(0)
    assert list(a) == list(b)
    for k in a:
        np.testing.assert_array_equal(a[k], b[k])
    n_params = sum(v.size for k, v in a.items() if "running" not in k)
    assert 11_000_000 < n_params < 13_000_000  # DBNet-R18 without the training-only thresh branch


def test_weight_packer_layouts():
    from pdf_table_b200 import weights

    rng = np.random.default_rng(0)
    w = rng.standard_normal((5, 24, 3, 3)).astype(np.float32)
    packed, bias = weights.pack_conv(w)
    assert packed.shape == (5, 9 * 32) and bias.shape == (256,)
    p = packed.reshape(5, 9, 32)
    np.testing.assert_array_equal(p[:, :, 24:], 0)
    np.testing.assert_array_equal(p[2, 4, :24], w[2, :, 1, 1].astype(np.float16))
    sw = rng.standard_normal((64, 3, 7, 7)).astype(np.float32)
    sp, _ = weights.pack_stem7x7(sw)
    sp = sp.reshape(64, 7, 8, 4)
    np.testing.assert_array_equal(sp[:, :, 7, :], 0)
    np.testing.assert_array_equal(sp[:, :, :, 3], 0)
    np.testing.assert_array_equal(sp[3, 2, 5, 1], sw[3, 1, 2, 5].astype(np.float16))
    dw = rng.standard_normal((16, 8, 2, 2)).astype(np.float32)
    dp, db = weights.pack_deconv2x2(dw, np.arange(8, dtype=np.float32))
    assert dp.shape == (32, 16)
    np.testing.assert_array_equal(dp[(1 * 2 + 0) * 8 + 3, :], dw[:, 3, 1, 0].astype(np.float16))
    np.testing.assert_array_equal(db[:32], np.tile(np.arange(8, dtype=np.float32), 4))
    blob = weights.pack_dbnet_r18(synth.dbnet_r18_state_dict(0))
    assert blob[:8] == b"DVWBLOB1" and len(blob) > 20_000_000


def test_convnextvit_oracle_matches_reference_golden():
    from oracle import convnextvit_ref

    g = np.load(os.path.join(GOLDEN, "convnextvit_seed0.npz"))
    n = int(g["n_crops"])
    crops = [g[f"crop{i}"] for i in range(n)]
    chunks = convnextvit_ref.preprocess(crops)
    np.testing.assert_array_equal(chunks[0].numpy(), g["chunk0"])
    np.testing.assert_allclose(chunks.double().sum(dim=(1, 2, 3)).numpy(), g["chunks_sum"], rtol=0, atol=1e-9)
    sd = synth.convnext_vit_state_dict(0)
    feats = convnextvit_ref.convnext_features(sd, chunks)
    np.testing.assert_allclose(feats[0].numpy(), g["feats0"], atol=2e-5, rtol=0)
    np.testing.assert_allclose(feats.abs().double().sum(dim=(1, 2, 3)).numpy(), g["feats_abs_sum"], rtol=1e-5)
    logits = convnextvit_ref.convnextvit_forward(sd, chunks)
    np.testing.assert_allclose(logits[:, :, ::32].numpy(), g["logits_sub"], atol=5e-5, rtol=0)
    np.testing.assert_array_equal(logits.argmax(-1).numpy(), g["argmax"])
    ids = convnextvit_ref.greedy_ids(logits)
    for i in range(n):
        np.testing.assert_array_equal(ids[i], g[f"ids{i}"])


def test_db_post_oracle_matches_reference_golden():
    """oracle/db_post_ref.py (restatement) == the reference's own DBPostProcess code run in the build container."""
    from oracle import db_post_ref
    from oracle.gen_golden_more import DB_POST_CASES

    g = np.load(os.path.join(GOLDEN, "db_post.npz"))
    for name, idx, h, w, n_lines, src_h, src_w in DB_POST_CASES:
        prob = synth.synthetic_prob_map(idx, h, w, n_lines)
        shape_list = np.array([src_h, src_w, h / float(src_h), w / float(src_w)])
        got = db_post_ref.db_postprocess(prob, shape_list, (src_h, src_w, 3))
        np.testing.assert_array_equal(got.astype(np.float32), g[name], err_msg=name)
        assert len(got) > 0


def test_dbnet_backend_post_oracle_matches_reference_golden():
    """oracle/db_post_ref.dbnet_postprocess == the reference's in-tree DBNet post-processor (model="db") run in the build container."""
    from oracle import db_post_ref
    from oracle.gen_golden_more import DBNET_POST_CASES

    g = np.load(os.path.join(GOLDEN, "dbnet_proc.npz"))
    for name, idx, h, w, n_lines, org_h, org_w in DBNET_POST_CASES:
        got = db_post_ref.dbnet_postprocess(synth.synthetic_prob_map(idx, h, w, n_lines), (org_h, org_w))
        np.testing.assert_array_equal(got, g["post_" + name])
        assert len(got) > 3


def test_min_area_rect_twin_is_bit_identical_to_cv2():
    """oracle/cv_geom_ref.min_area_rect / box_points -- the numpy twin of the device code in csrc/db_post.cu -- against the cv2
    of this image, every bit of (centre, size, angle) and of the four box points: random filled quads and discs through
    findContours, plus the degenerate hulls (1 and 2 points, collinear points)."""
    import sys

    import cv2

    sys.path.insert(0, os.path.join(os.path.dirname(GOLDEN), "..", "tools"))
    import min_area_rect_probe as probe
    from oracle import cv_geom_ref as G

    n = 0
    for c in probe.contours(300, seed=3):
        r = cv2.minAreaRect(c)
        m = G.min_area_rect([(int(p[0]), int(p[1])) for p in c.reshape(-1, 2)])
        got = np.array([m[0][0], m[0][1], m[1][0], m[1][1], m[2]], np.float32)
        want = np.array([r[0][0], r[0][1], r[1][0], r[1][1], r[2]], np.float32)
        assert (got.view(np.uint32) == want.view(np.uint32)).all(), (c.reshape(-1, 2).tolist(), got, want)
        assert (G.box_points(m).view(np.uint32) == cv2.boxPoints(r).view(np.uint32)).all()
        n += 1
    assert n >= 300
    for pts in ([[3, 4], [10, 9]], [[3, 4], [3, 9]], [[3, 4], [9, 4]], [[5, 5]], [[3, 4], [10, 2]], [[3, 4], [5, 6], [7, 8]],
                [[0, 0], [4, 0], [4, 1], [0, 1]]):
        r = cv2.minAreaRect(np.array(pts, np.int32).reshape(-1, 1, 2))
        m = G.min_area_rect([tuple(p) for p in pts])
        assert (r[0][0], r[0][1], r[1][0], r[1][1], r[2]) == (float(m[0][0]), float(m[0][1]), float(m[1][0]), float(m[1][1]), float(m[2])), pts


def test_clipper_offset_known_answers():
    """Round-join offset of an axis-aligned rectangle: every point lies within 1 of distance d from the source
    rectangle, the axis extremes are exactly +-d, and tiny deltas use the reduced arc tolerance."""
    from oracle.db_post_ref import clipper_offset_round

    rect = [(10, 20), (110, 20), (110, 50), (10, 50)]
    cround = lambda v: int(v - 0.5) if v < 0 else int(v + 0.5)  # clipper.cpp Round()
    for d in (0.7, 3.0, 9.5, 40.0):
        pts = np.array(clipper_offset_round(rect, d))
        assert pts[:, 0].min() == cround(10 - d) and pts[:, 0].max() == cround(110 + d)
        assert pts[:, 1].min() == cround(20 - d) and pts[:, 1].max() == cround(50 + d)
        dx = np.maximum(np.maximum(10 - pts[:, 0], pts[:, 0] - 110), 0)
        dy = np.maximum(np.maximum(20 - pts[:, 1], pts[:, 1] - 50), 0)
        assert (np.abs(np.hypot(dx, dy) - d) <= 0.75).all()
    # orientation: the reversed path gives the same point set
    a = set(clipper_offset_round(rect, 5.0))
    b = set(clipper_offset_round(rect[::-1], 5.0))
    assert a == b


def test_packed_dbnet_blob_computes_the_reference_function():
    """The host half of the weight path on CPU: pack_dbnet_r18 (BatchNorm folding, K-major tap order, channel padding, the 7x7
    stem and 2x2 transposed-conv layouts) -> write_blob -> parsed back with the container format csrc/capi.cu reads -> the
    network run FROM THE PACKED TENSORS (oracle/blob_ref.py) equals the oracle forward up to the fp16 rounding of the weights."""
    from oracle import blob_ref
    from pdf_table_b200 import weights

    sd = synth.dbnet_r18_state_dict(0)
    blob = weights.pack_dbnet_r18(sd)
    t = blob_ref.read_blob(blob)
    assert t["stem.w"].dtype == np.float16 and t["stem.w"].shape == (64, 7 * 32) and t["layer2.0.down.w"].shape == (128, 64)
    assert t["bin.deconv1.w"].shape == (256, 64) and t["bin.deconv2.w"].shape == (64, 4)
    x = torch.from_numpy(np.random.default_rng(41).standard_normal((2, 3, 96, 128)).astype(np.float32))
    want = dbnet_ref.dbnet_r18_forward(sd, x)
    got = blob_ref.dbnet_r18_from_blob(t, x)
    err = float((got - want).abs().max())
    assert got.shape == want.shape == (2, 1, 96, 128) and err < 4e-3, err  # measured 2.4e-3: the fp16 rounding of the weights
    # the check discriminates: the same tensors with two taps of one 3x3 layer swapped are an order of magnitude further away
    bad = dict(t)
    w = t["layer1.0.conv1.w"].reshape(64, 9, 64).copy()
    w[:, [0, 8]] = w[:, [8, 0]]
    bad["layer1.0.conv1.w"] = w.reshape(64, 9 * 64)
    assert float((blob_ref.dbnet_r18_from_blob(bad, x) - want).abs().max()) > 10 * err
    with pytest.raises(ValueError):
        blob_ref.read_blob(blob[:4096])


def test_packed_convnextvit_blob_computes_the_reference_function():
    """pack_convnext_vit on CPU: the packed tensors (layer_scale and attention scale folded, q | k | v concatenated, the (2,1)
    down-sampling conv and the projections K-major, position table without the CLS row) turned back into a reference-keyed
    state_dict (oracle/blob_ref.py) give the oracle's logits up to the fp16 rounding of the GEMM weights -- which is most of
    the engine's own distance from the fp32 oracle (6e-3 of its 8.5e-3)."""
    from oracle import blob_ref
    from oracle import convnextvit_ref as ref
    from pdf_table_b200 import weights

    sd = synth.convnext_vit_state_dict(0)
    t = blob_ref.read_blob(weights.pack_convnext_vit(sd))
    sd2 = blob_ref.convnext_vit_state_dict_from_blob(t)
    assert set(sd) - set(sd2) <= {"cnn_model.layernorm.bias", "cnn_model.layernorm.weight", "vitstr.vit.embeddings.cls_token"}  # unused by the forward
    chunks = ref.preprocess([synth.synthetic_text_crop(700 + i, 32, 320) for i in range(2)])
    want = ref.convnextvit_forward(sd, chunks)
    got = ref.convnextvit_forward(sd2, chunks)
    err = float((got - want).abs().max())
    assert err < 1.2e-2, err  # measured 6.1e-3 on logits with sigma 2.06
    assert float((got.argmax(-1) == want.argmax(-1)).float().mean()) >= 0.999
    # discriminates: forgetting to undo the folded attention scale moves the logits by far more
    bad = dict(sd2)
    k = "vitstr.vit.encoder.layer.0.attention.attention.query.weight"
    bad[k] = sd2[k] / np.float32(8.0)
    assert float((ref.convnextvit_forward(bad, chunks) - want).abs().max()) > 5 * err


def test_crnn_oracle_matches_the_reference_module():
    """oracle/crnn_ref.py (conv stack, hand-rolled bidirectional LSTMs, embeddings, classifier) against the logits the reference
    CRNN module produced in the build container on the same seeded weights (oracle/gen_golden_crnn.py)."""
    import torch

    from oracle import crnn_ref
    from oracle.gen_golden_crnn import LABELS, case_input
    from pdf_table_b200 import synth, weights

    sd = synth.crnn_state_dict(0, LABELS)
    g = np.load(os.path.join(GOLDEN, "crnn_seed0.npz"))
    for n, w in ((2, 300), (1, 640), (3, 64)):
        got = crnn_ref.crnn_forward(sd, torch.from_numpy(case_input(n, w))).numpy()
        want = g[f"logits_{n}x{w}"]
        assert got.shape == want.shape == (n, w // 4, LABELS)
        assert float(np.abs(got - want).max()) < 1e-5
    assert len(weights.pack_crnn(sd)) > 0


def test_cell_text_matching_oracle_equals_the_reference_functions():
    """oracle/match_ref.py against the indices the reference's own find_top1_mach_box / box_in_other_box / distance /
    compute_iou_v2 produced in the build container (their source executed as is: oracle/gen_golden_match.py)."""
    from oracle import match_ref

    g = np.load(os.path.join(GOLDEN, "match_seed0.npz"))
    n = 0
    for i in range(6):
        got = match_ref.match(g[f"texts{i}"], g[f"cells{i}"])
        assert got == g[f"top1_{i}"].tolist()
        n += len(got)
    assert n > 200
    # known answers: containment beats IoU, the FIRST containing cell wins, ties go to the first cell
    cells = [[0, 0, 100, 50], [0, 0, 100, 50], [100, 0, 200, 50]]
    assert match_ref.match([[10, 10, 40, 30]], cells) == [0]
    assert match_ref.match([[90, 10, 160, 30]], cells) == [2]
    assert match_ref.match([[300, 300, 320, 310]], cells) == [2]
