"""SURVEY.md a2 on the GPU: the PP-OCRv4 mobile detector (PPLCNetV3-0.75 + RSE-FPN + DBHead) through dv_dbnet_forward /
dv_dbnet_forward_u8 on a "pp_det" handle against the fp32 oracle of the published architecture (oracle/pp_det_ref.py -- parity
unpinned against the hub ONNX, see its header), and through OcrDetectionTask(model="db_pp", backbone="PPLCNetV3") against
the explicit composition a1 -> a2 -> a3.

Tolerance (fp16 operands, fp32 accumulation, fp16 activations through ~45 layers): |dprob| <= PROB_TOL on the sigmoid output."""
import numpy as np
import pytest
import torch

from oracle import pp_det_ref
from pdf_table_b200 import pp_det_graph, predictors, synth
from pdf_table_b200.engine import Engine

pytestmark = pytest.mark.gpu
PROB_TOL = 1e-2
MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)


@pytest.fixture(scope="module")
def det():
    sd = synth.pp_ocrv4_det_state_dict(0)
    eng = Engine("pp_det", pp_det_graph.pack_pp_det(sd))
    yield eng, sd
    eng.close()


def test_pp_det_network_vs_oracle(det):
    eng, sd = det
    rng = np.random.default_rng(11)
    for n, h, w in ((2, 96, 160), (1, 256, 320), (3, 64, 64)):
        x = torch.from_numpy(rng.standard_normal((n, 3, h, w)).astype(np.float32))
        want, fuse = pp_det_ref.pp_det_forward(sd, x, return_fuse=True)
        got = eng.dbnet_forward(x.cuda()).cpu()
        assert tuple(got.shape) == (n, 1, h, w)
        err = float((got - want).abs().max())
        print(f"pp_det {n}x{h}x{w}: max |dprob| = {err:.3e} (prob std {float(want.std()):.2f})")
        assert err <= PROB_TOL
        _, meta = pp_det_graph.build_pp_det(sd)
        f = eng.debug_tensor(f"t{meta['fuse']}").cpu()
        ferr = float((f - fuse).abs().max()) / float(fuse.abs().max())
        print(f"  fused neck map: rel max|err| = {ferr:.3e}")
        assert ferr < 1e-2


def test_pp_det_u8_path_equals_fp32_path(det):
    """The fused flip / normalise of the stem kernel is the numpy expression of PPOcrDetectionPreprocessor (bit-exact)."""
    eng, _ = det
    page = synth.synthetic_page(3, 160, 224)
    mean, std = np.array(MEAN, np.float32).reshape(1, 1, 3), np.array(STD, np.float32).reshape(1, 1, 3)
    x = ((page[:, :, ::-1].astype("float32") * np.float32(1.0 / 255.0) - mean) / std).transpose(2, 0, 1)[None]
    a = eng.dbnet_forward(torch.from_numpy(np.ascontiguousarray(x)).cuda())
    b = eng.dbnet_forward_u8(torch.from_numpy(page[None]).cuda(), MEAN, STD, 1.0 / 255.0, True)
    assert torch.equal(a, b)


def test_detection_task_with_the_pp_ocrv4_backbone(det):
    _, sd = det
    task = predictors.OcrDetectionTask(model="db_pp", backbone="PPLCNetV3", state_dict=sd)
    pages = [synth.synthetic_page(s, 300, 500) for s in (1, 2)]
    res = task(pages)
    assert isinstance(res, list) and len(res) == 2
    for r in res:
        assert r.dtype == np.float32 and r.ndim == 2 and r.shape[1] == 8
    page, _ = predictors.det_resize_for_test(pages[0], 960, "max")
    prob = task.predictor.dbnet_forward_u8(torch.from_numpy(page[None].copy()).cuda(), MEAN, STD, 1 / 255.0, True)
    boxes, counts = task.predictor.db_boxes(prob, [(300, 500)], 0.2, 0.6, 1.5, 1000)
    np.testing.assert_array_equal(boxes.cpu().numpy()[0, : int(counts[0])], res[0])
    with pytest.raises(RuntimeError):
        predictors.OcrDetectionTask(model="db", backbone="PPLCNetV3", state_dict=sd)


def test_pp_det_fp32x_meets_the_north_star_tolerance():
    """precision="fp32x" on the graph executor (fp32 buffers, split-fp16 GEMMs / 3x3 convs / transposed conv): the probability map
    within 1e-3 of the fp32 oracle, and the boxes of the task equal to the DB post-process of the oracle's own map."""
    from oracle import db_post_ref

    sd = synth.pp_ocrv4_det_state_dict(0)
    eng = Engine("pp_det", pp_det_graph.pack_pp_det(sd, precise=True))
    rng = np.random.default_rng(11)
    for n, h, w in ((2, 96, 160), (1, 256, 320)):
        x = torch.from_numpy(rng.standard_normal((n, 3, h, w)).astype(np.float32))
        want = pp_det_ref.pp_det_forward(sd, x)
        err = float((eng.dbnet_forward(x.cuda()).cpu() - want).abs().max())
        print(f"pp_det fp32x {n}x{h}x{w}: max |dprob| = {err:.3e}")
        assert err <= 1e-3
    eng.close()
    task = predictors.OcrDetectionTask(model="db_pp", backbone="PPLCNetV3", state_dict=sd, precision="fp32x")
    assert len(task([synth.synthetic_page(1, 300, 500)])) == 1
