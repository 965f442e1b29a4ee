"""db_boxes (threshold + box seed) vs the reference-generated golden boxes and the CPU oracle.

Parity statement (DESIGN.md "DB post-process"): contour discovery, border tracing, convex hull, cv2.minAreaRect (OpenCV
4.13: counter-clockwise hull, cross-product edge selection, double angle normalisation -- bit-identical to cv2 on every
probed contour, tests/test_oracle_cpu.py), boxPoints, the fillPoly mask, the Clipper offset and all integer box arithmetic
are restated exactly.  The tests therefore require IDENTICAL integer output: same number of boxes, same order, every
corner equal."""
import os

import numpy as np
import pytest
import torch

from oracle import db_post_ref
from oracle.gen_golden_more import DB_POST_CASES
from pdf_table_b200 import synth

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _run(eng, prob, src_h, src_w, **kw):
    boxes, counts, ovf = eng.db_boxes(torch.from_numpy(prob)[None, None].cuda(), [(src_h, src_w)], check_overflow=True, **kw)
    eng.sync()
    assert ovf == 0
    n = int(counts.cpu()[0])
    return boxes.cpu().numpy()[0, :n]


def _compare(got, want, name):
    assert got.shape == want.shape, f"{name}: {got.shape[0]} boxes vs {want.shape[0]}"
    if len(want) == 0:
        return 1.0
    dev = np.abs(got - want).max(axis=1)
    assert dev.max() == 0.0, f"{name}: corner deviation {dev.max()} px in {int((dev > 0).sum())} of {len(dev)} boxes\n{got[dev.argmax()]}\n{want[dev.argmax()]}"
    return 1.0


def test_db_boxes_reference_golden(post_engine):
    g = np.load(os.path.join(GOLDEN, "db_post.npz"))
    fracs = []
    for name, idx, h, w, n_lines, src_h, src_w in DB_POST_CASES:
        prob = synth.synthetic_prob_map(idx, h, w, n_lines)
        got = _run(post_engine, prob, src_h, src_w)
        fracs.append((name, len(got), _compare(got, g[name], name)))
    print("db_boxes vs reference golden (name, boxes, identical fraction):", fracs)
    tot = sum(n for _, n, _ in fracs)
    ident = sum(n * f for _, n, f in fracs)
    assert ident == tot


def test_db_boxes_dbnet_backend_reference_golden(post_engine):
    """model="db": dv_db_boxes_dbnet vs the reference's in-tree DBNet post-processor (tests/golden/dbnet_proc.npz)."""
    from oracle.gen_golden_more import DBNET_POST_CASES

    g = np.load(os.path.join(GOLDEN, "dbnet_proc.npz"))
    for name, idx, h, w, n_lines, org_h, org_w in DBNET_POST_CASES:
        prob = synth.synthetic_prob_map(idx, h, w, n_lines)
        got = _run(post_engine, prob, org_h, org_w, box_thresh=0.3, variant="db")
        want = g["post_" + name]
        assert got.shape == want.shape, f"{name}: {got.shape[0]} boxes vs {want.shape[0]}"
        np.testing.assert_array_equal(got.astype(np.int64), want, err_msg=name)
        np.testing.assert_array_equal(got.astype(np.int64), db_post_ref.dbnet_postprocess(prob, (org_h, org_w)))


def test_db_boxes_vs_oracle_batch(post_engine):
    """A batch of 4 different pages in one call (page indexing, per-page src sizes) against the oracle."""
    probs = [synth.synthetic_prob_map(10 + i, 480, 640, 25) for i in range(4)]
    src = [(480, 640), (960, 1280), (500, 700), (480, 640)]
    boxes, counts = post_engine.db_boxes(torch.from_numpy(np.stack(probs))[:, None].cuda(), src)
    post_engine.sync()
    boxes, counts = boxes.cpu().numpy(), counts.cpu().numpy()
    tot = ident = 0
    for i in range(4):
        want = db_post_ref.db_postprocess(probs[i], np.array([src[i][0], src[i][1], 480 / src[i][0], 640 / src[i][1]]), src[i])
        f = _compare(boxes[i, : counts[i]], want.astype(np.float32).reshape(-1, 8), f"page{i}")
        tot += len(want)
        ident += f * len(want)
    print(f"db_boxes batch: {tot} boxes, identical fraction {ident / tot:.3f}")
    assert ident == tot and tot > 40


def test_db_boxes_edge_cases(post_engine):
    # empty map -> no boxes; full map -> one page-sized component touching the frame
    z = np.zeros((64, 96), np.float32)
    assert len(_run(post_engine, z, 64, 96)) == 0
    f = np.full((64, 96), 0.9, np.float32)
    want = db_post_ref.db_postprocess(f, np.array([64, 96, 1.0, 1.0]), (64, 96))
    got = _run(post_engine, f, 64, 96)
    _compare(got, want.astype(np.float32).reshape(-1, 8), "full")
    # max_candidates smaller than the number of contours: the reference keeps the FIRST contours in cv2 order
    prob = synth.synthetic_prob_map(0, 320, 480, 12)
    want = db_post_ref.db_postprocess(prob, np.array([320, 480, 1.0, 1.0]), (320, 480), max_candidates=5)
    got = _run(post_engine, prob, 320, 480, max_candidates=5)
    _compare(got, want.astype(np.float32).reshape(-1, 8), "max_candidates")
    # thresholds
    want = db_post_ref.db_postprocess(prob, np.array([320, 480, 1.0, 1.0]), (320, 480), thresh=0.3, box_thresh=0.7, unclip_ratio=2.0)
    got = _run(post_engine, prob, 320, 480, thresh=0.3, box_thresh=0.7, unclip_ratio=2.0)
    _compare(got, want.astype(np.float32).reshape(-1, 8), "thresholds")
