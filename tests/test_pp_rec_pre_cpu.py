"""the PP-OCR recogniser pre-process.  CPU side: the oracle restatement and the host batching rule against the
golden batches produced by the reference's own PPOcrRecPreProcessor (oracle/gen_golden_pp_rec_pre.py)."""
import os

import numpy as np

from oracle import gen_golden_pp_rec_pre as gen
from oracle import pp_rec_pre_ref as ref
from pdf_table_b200 import predictors

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pp_rec_pre.npz")


def test_oracle_matches_reference_golden():
    g = np.load(GOLDEN)
    batches = ref.preprocess([c if c.ndim == 3 else np.stack([c] * 3, -1) for c in gen.crops()])
    assert len(batches) == int(g["n_batches"])
    assert np.array_equal(batches[0]["indices"], g["indices"])
    for k, b in enumerate(batches):
        assert b["batch_beg_img_no"] == int(g[f"beg{k}"])
        assert b["image"].dtype == np.float32 and b["image"].shape == g[f"image{k}"].shape
        assert np.array_equal(b["image"], g[f"image{k}"])  # bit-exact


def test_host_batch_plan_matches_oracle_and_golden():
    g = np.load(GOLDEN)
    shapes = [c.shape[:2] for c in gen.crops()]
    idx, plan = predictors.pp_rec_batch_plan(shapes)
    ridx, rplan = ref.batch_plan(shapes)
    assert np.array_equal(idx, ridx) and np.array_equal(idx, g["indices"])
    assert plan == rplan
    for k, (beg, img_w, widths) in enumerate(plan):
        assert beg == int(g[f"beg{k}"]) and img_w == g[f"image{k}"].shape[3] and len(widths) == g[f"image{k}"].shape[0]
    # the clamps: a 30:1 crop is cut to limited_max_width, a 0.27:1 crop is widened to limited_min_width
    assert plan[-1][1] == 1280 and plan[-1][2] == [1280]
    assert min(plan[0][2]) == 16


def test_host_batch_plan_edge_cases():
    assert predictors.pp_rec_batch_plan([])[1] == []
    idx, plan = predictors.pp_rec_batch_plan([(48, 320)])
    assert list(idx) == [0] and plan == [(0, 320, [320])]
    # seven equal crops: two batches, the second of one
    idx, plan = predictors.pp_rec_batch_plan([(24, 100)] * 7)
    assert [len(p[2]) for p in plan] == [6, 1] and all(p[1] == 320 for p in plan)
