"""SURVEY.md a4 on the GPU: PPOcrRecPreProcessor (host cv2.resize + dv_pp_rec_normalise) against the golden batches of the
reference's own class -- bit-exact."""
import os

import numpy as np
import pytest
import torch

from oracle import gen_golden_pp_rec_pre as gen
from pdf_table_b200 import predictors
from pdf_table_b200.engine import Engine

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pp_rec_pre.npz")


@pytest.fixture(scope="module")
def pre():
    eng = Engine("post")
    yield predictors.PPOcrRecPreProcessor(eng)
    eng.close()


def test_pp_rec_preprocess_matches_reference_golden(pre):
    g = np.load(GOLDEN)
    batches = pre(gen.crops())
    assert len(batches) == int(g["n_batches"])
    for k, b in enumerate(batches):
        assert np.array_equal(b["indices"], g["indices"]) and b["batch_beg_img_no"] == int(g[f"beg{k}"])
        img = b["image"]
        assert img.is_cuda and img.dtype == torch.float32
        assert np.array_equal(img.cpu().numpy(), g[f"image{k}"])  # bit-exact, padding included


def test_pp_rec_preprocess_single_input_and_bad_type(pre):
    crop = gen.crops()[0]
    (b,) = pre(crop)  # a bare ndarray is wrapped into a list, as in the reference
    assert tuple(b["image"].shape) == (1, 3, 48, 480) and b["batch_beg_img_no"] == 0
    with pytest.raises(TypeError):
        pre([3.14])


def test_pp_rec_normalise_rejects_bad_arguments():
    eng = Engine("post")
    with pytest.raises(ValueError):
        eng.pp_rec_normalise(torch.zeros((2, 48, 64, 4), dtype=torch.uint8, device="cuda"), torch.zeros(2, dtype=torch.int32, device="cuda"))
    eng.close()
