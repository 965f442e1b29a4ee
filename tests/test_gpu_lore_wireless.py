"""Lore `wireless` configuration on the engine: the ResNet-18 key-point detector (model kind "lore_resnet18", csrc/lore_net.cu
build_r18) vs oracle/lore_wireless_ref.py (pinned to the reference's LoreDetectModel by tests/golden/lore_resnet18_seed0.npz),
its sparse ax / cr evaluation, and OcrTableStructureTask(task_type="wireless") end to end in the upper-left-anchored frame."""
import os

import numpy as np
import pytest
import torch

from oracle import lore_decode_ref, lore_processor_ref, lore_wireless_ref
from pdf_table_b200 import predictors, synth, weights
from pdf_table_b200.engine import Engine

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
# fp16 operands / fp32 accumulation through ~45 conv layers, relative to max(1, max|oracle|) of the tensor; measured 1.1e-3 ..
# 1.8e-3 on the regression heads and 4.1e-3 on the sigmoid-ed heat map (its logits reach +-10 with the synthetic weights)
REL_TOL = 6e-3
LEVELS = ("x0", "x1", "x2", "x3", "x4", "x3_", "x2_", "x1_", "feat")


def _unpack(maps):
    m = maps.cpu().numpy()
    return {"hm": m[..., 0:2], "reg": m[..., 2:4], "wh": m[..., 4:12], "st": m[..., 12:20]}


def _head_errors(got, want_of):
    errs = {}
    for k in ("hm", "reg", "wh", "st"):
        want = want_of(k).transpose(0, 2, 3, 1)
        if k == "hm":
            want = 1.0 / (1.0 + np.exp(-want))
        errs[k] = float(np.abs(got[k] - want).max()) / max(1.0, float(np.abs(want).max()))
    return errs


def _cell_feature_error(eng, post, maps, out, n, hw, cells=30, seed=5):
    """Plants peaks (random-weight heat maps have none above the gates), decodes them and compares the sparse ax / cr evaluation
    with the oracle's dense head maps gathered at the same points."""
    h, w = hw
    planted = maps.clone()
    prng = np.random.default_rng(seed)
    for i in range(n):
        for _ in range(cells):
            planted[i, int(prng.integers(3, h - 3)), int(prng.integers(3, w - 3)), 0] = float(prng.uniform(0.5, 0.95))
    eye = np.tile(np.array([[1.0, 0, 0], [0, 1.0, 0]]), (n, 1, 1))
    dec = post.lore_decode(planted, None, None, None, eye, wiz_rev=False, vis_thresh=0.3)
    feat, offsets = eng.lore_cell_features(dec, max_rows=4096, check_overflow=True)
    counts, offs = dec["counts"].cpu().numpy(), offsets.cpu().numpy()
    worst, scale = 0.0, 1.0
    for i in range(n):
        ax, cr = out["ax"][i].numpy().reshape(256, -1), out["cr"][i].numpy().reshape(256, -1)
        a_idx, c_idx = dec["ax_idx"].cpu().numpy()[i, : counts[i]], dec["cr_idx"].cpu().numpy()[i, : counts[i]]
        want = ax[:, a_idx].T + sum(cr[:, c_idx[:, k]].T for k in range(4))
        worst = max(worst, float(np.abs(feat.cpu().numpy()[offs[i]: offs[i + 1]] - want).max()))
        scale = max(scale, float(np.abs(want).max()))
    assert counts.sum() >= cells and offs[-1] == counts.sum()
    return worst, scale, int(counts.sum())


@pytest.mark.parametrize("precise,tol", [(False, REL_TOL), (True, 1e-3)])
def test_wireless_detector_vs_reference_golden_and_oracle(post_engine, precise, tol):
    """Head maps against the reference module's own outputs; every stage / top-down level, the stride-4 feature map and the
    sparse ax / cr features against the oracle on a 2-image non-square batch.  fp32x (split-fp16 operand pairs): 1e-3."""
    sd = synth.lore_resnet18_state_dict(0)
    eng = Engine("lore_resnet18", weights.pack_lore_resnet18(sd, precise=precise))
    g = np.load(os.path.join(GOLDEN, "lore_resnet18_seed0.npz"))
    errs = _head_errors(_unpack(eng.lore_detect_forward(torch.from_numpy(g["x"]).cuda())), lambda k: g[k])
    print(f"wireless detector (precise={precise}) relative max|err| vs reference golden:", errs)
    assert max(errs.values()) < tol
    rng = np.random.default_rng(23)
    x = torch.from_numpy(rng.standard_normal((2, 3, 128, 192)).astype(np.float32))
    maps = eng.lore_detect_forward(x.cuda())
    out = lore_wireless_ref.lore_resnet18_forward(sd, x)
    for name in LEVELS:
        want = out[name].numpy()
        err = float(np.abs(eng.debug_tensor(name).cpu().numpy() - want).max())
        print(f"  {name}: max|err| {err:.3e} (max|x| {float(np.abs(want).max()):.2f})")
        assert err < tol * max(float(np.abs(want).max()), 1.0), name
    errs = _head_errors(_unpack(maps), lambda k: out[k].numpy())
    print("  heads:", errs)
    assert max(errs.values()) < tol
    worst, scale, n_cells = _cell_feature_error(eng, post_engine, maps, out, 2, (32, 48))
    print(f"  cell features: {n_cells} cells, max|err| {worst:.3e} (max|x| {scale:.2f})")
    assert worst < (2 if not precise else 1) * tol * scale  # sum of five head evaluations
    eng.close()


def test_wireless_u8_input_equals_fp32_input():
    """The fused normalisation of the 4-channel stem layout is bit-exact w.r.t. numpy's float64 expression."""
    sd = synth.lore_resnet18_state_dict(0)
    eng = Engine("lore_resnet18", weights.pack_lore_resnet18(sd))
    page = synth.synthetic_page(4, 128, 192)
    mean = np.array(eng.LORE_MEAN, dtype=np.float32).reshape(1, 1, 3)
    std = np.array(eng.LORE_STD, dtype=np.float32).reshape(1, 1, 3)
    x = ((page / 255. - mean) / std).astype(np.float32).transpose(2, 0, 1)[None]
    a = eng.lore_detect_forward(torch.from_numpy(x).cuda())
    b = eng.lore_detect_forward_u8(torch.from_numpy(page[None]).cuda())
    assert torch.equal(a, b)
    with pytest.raises(Exception):
        eng.lore_detect_forward(torch.zeros((1, 3, 96, 128), device="cuda"))  # H, W must be multiples of 64
    eng.close()


def test_table_structure_task_wireless():
    """OcrTableStructureTask(model="Lore", task_type="wireless"): 768 x 768 upper-left frame, no corner snapping, 2-D position
    embeddings; the polygons equal the oracle decode of the engine's own maps bit for bit and the logical locations follow the
    processor oracle called with dets=."""
    sd = synth.lore_resnet18_state_dict(0)
    sd["hm.8.bias"] = np.array([-0.6, -3.5], np.float32)  # random weights: shift the heat map so that cells pass the 0.2 gate
    psd = synth.lore_processor_state_dict(0)
    task = predictors.OcrTableStructureTask(model="Lore", task_type="wireless", state_dict=(sd, psd))
    assert task.resolution == (768, 768) and task.upper_left and task.wiz_2dpe and not task.wiz_rev and task.predictor.kind == "lore_resnet18"
    pages = [synth.synthetic_page(7, 700, 900), synth.synthetic_page(8, 1024, 768)]
    res = task(pages)
    assert len(res) == 2
    for r in res:
        assert r["polygons"].dtype == np.float32 and r["polygons"].shape[1] == 8 and r["logi"].shape == (len(r["polygons"]), 4)
        assert np.array_equal(r["logi"], np.floor(r["logi"]))
    assert sum(len(r["polygons"]) for r in res) > 5
    eng, post, proc = task.predictor, task.post, task.processor
    pre = [predictors.lore_preprocess(p, (768, 768), upper_left=True) for p in pages]
    maps = eng.lore_detect_forward_u8(torch.from_numpy(np.stack([w for w, _ in pre])).cuda())
    m = maps.cpu().numpy()
    for i, (_, meta) in enumerate(pre):
        assert list(meta[:2]) == [0, 0]
        z = np.zeros((1, 192, 192), np.float32)
        want = lore_decode_ref.lore_decode(m[i, :, :, 0:2].transpose(2, 0, 1), m[i, :, :, 2:4].transpose(2, 0, 1), m[i, :, :, 4:12].transpose(2, 0, 1),
                                           m[i, :, :, 12:20].transpose(2, 0, 1), z, z, meta, upper_left=True, wiz_rev=False, vis_thresh=0.2)
        np.testing.assert_array_equal(res[i]["polygons"], want["polygons"])
    inv = np.stack([predictors.lore_affine_upper_left([np.float32(mm[0]), np.float32(mm[1])], np.float32(mm[2]), 192, 192, True) for _, mm in pre])
    dec = post.lore_decode(maps, None, None, None, inv, wiz_rev=False, vis_thresh=0.2)
    feat, offsets = eng.lore_cell_features(dec, max_rows=6000, check_overflow=True)
    offs, counts = offsets.cpu().numpy(), dec["counts"].cpu().numpy()
    for i in range(2):
        f = feat[offs[i]: offs[i + 1]].cpu()
        if len(f) == 0:
            continue
        dets = dec["dets_feat"][i, : counts[i]].cpu().to(torch.int64)
        _, stacked = lore_processor_ref.lore_processor_forward(psd, f, dets=dets)
        want_logi = lore_decode_ref.round_logic(stacked.numpy())
        safe = np.abs((stacked.numpy() - np.floor(stacked.numpy())) - 0.5) > 2e-3
        np.testing.assert_array_equal(res[i]["logi"][safe], want_logi[safe])
        assert safe.mean() > 0.98
    # host-warp variant and the device-cut tables of the orchestrator agree with the per-image call
    host = predictors.OcrTableStructureTask(model="Lore", task_type="wireless", state_dict=(sd, psd), host_warp=True)(pages[:1])
    np.testing.assert_array_equal(host[0]["polygons"], res[0]["polygons"])
    tables = task.recognize_tables(pages[0], [{"bbox": [0, 0, 900, 700]}])
    np.testing.assert_array_equal(tables[0][1]["polygons"], res[0]["polygons"])
