"""Lore detector / cell features / processor on the engine vs the oracle restatements (which are pinned against the
reference modules by tests/test_lore_oracle_cpu.py) and the reference-generated golden fixtures."""
import os

import numpy as np
import pytest
import torch

from oracle import lore_decode_ref, lore_net_ref, lore_processor_ref
from pdf_table_b200 import synth, weights
from pdf_table_b200.engine import Engine

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")

# fp16 operands / fp32 accumulation through ~60 conv layers (16 of them deformable, whose sampling positions depend
# on the activations): tolerance on max|err| RELATIVE to max(1, max|oracle|) of the tensor; measured 1.2e-3 .. 1.6e-3.
REL_TOL = 4e-3
PROC_TOL = 2e-4  # split-fp16 GEMMs: ~fp32 accuracy on outputs of magnitude up to ~10


@pytest.fixture(scope="module")
def lore_engine():
    eng = Engine("lore_dla34", weights.pack_lore_dla34(synth.lore_dla34_state_dict(0)))
    yield eng
    eng.close()


@pytest.fixture(scope="module")
def proc_engine():
    eng = Engine("lore_processor", weights.pack_lore_processor(synth.lore_processor_state_dict(0)))
    yield eng
    eng.close()


def _unpack(maps):
    m = maps.cpu().numpy()
    return {"hm": m[..., 0:2], "reg": m[..., 2:4], "wh": m[..., 4:12], "st": m[..., 12:20]}


def test_lore_detector_reference_golden(lore_engine):
    g = np.load(os.path.join(GOLDEN, "lore_dla34_seed0.npz"))
    maps = lore_engine.lore_detect_forward(torch.from_numpy(g["x"]).cuda())
    lore_engine.sync()
    got = _unpack(maps)
    errs = {}
    for k in ("hm", "reg", "wh", "st"):
        want = g[k].transpose(0, 2, 3, 1)
        if k == "hm":
            want = 1.0 / (1.0 + np.exp(-want))
        errs[k] = float(np.abs(got[k] - want).max()) / max(1.0, float(np.abs(want).max()))
    feat = lore_engine.debug_tensor("feat").cpu().numpy()
    print("lore detector relative max|err| vs reference golden:", errs)
    assert max(errs.values()) < REL_TOL
    assert np.isfinite(feat).all()


def test_lore_detector_levels_vs_oracle(lore_engine):
    """Per-level parity on a 2-image batch with non-square input: localises an error to a DLA level / the neck."""
    sd = synth.lore_dla34_state_dict(0)
    rng = np.random.default_rng(21)
    x = torch.from_numpy(rng.standard_normal((2, 3, 96, 160)).astype(np.float32))
    lore_engine.lore_detect_forward(x.cuda())
    lore_engine.sync()
    base = lore_net_ref.dla34_base(sd, x)
    for lvl in range(0, 6):
        got = lore_engine.debug_tensor(f"level{lvl}").cpu().numpy()
        want = base[lvl].numpy()
        err = float(np.abs(got - want).max())
        scale = float(np.abs(want).max())
        print(f"level{lvl}: max|err| {err:.3e} (max|x| {scale:.2f})")
        assert err < REL_TOL * max(scale, 1.0), f"level{lvl}"
    want = lore_net_ref.lore_dla34_features(sd, x).numpy()
    got = lore_engine.debug_tensor("feat").cpu().numpy()
    err = float(np.abs(got - want).max())
    print(f"feat: max|err| {err:.3e} (max|x| {float(np.abs(want).max()):.2f})")
    assert err < REL_TOL * max(float(np.abs(want).max()), 1.0)


def test_fused_dcn_equals_the_three_launch_path(lore_engine, monkeypatch):
    """dcn_fused_tcgen05 (sampling producers -> shared-memory A tiles -> tcgen05 GEMM, no column buffer) against k_dcn_im2col +
    the flat GEMM it replaces (DV_DCN_FUSED=0): same fp16 blend weights, same K order -> bit-identical head maps and feature
    map, on a ragged map size (24 x 40: partial 8 x 16 tiles on the right edge) and on a batch whose maps are tile multiples."""
    ref_eng = Engine("lore_dla34", weights.pack_lore_dla34(synth.lore_dla34_state_dict(0)))
    rng = np.random.default_rng(77)
    for shape in ((2, 3, 96, 160), (1, 3, 256, 320), (3, 3, 64, 64)):
        x = torch.from_numpy(rng.standard_normal(shape).astype(np.float32)).cuda()
        monkeypatch.setenv("DV_DCN_FUSED", "0")  # read when the network is planned for a new input shape
        want = ref_eng.lore_detect_forward(x).clone()
        monkeypatch.delenv("DV_DCN_FUSED")
        want_feat = ref_eng.debug_tensor("feat").clone()
        got = lore_engine.lore_detect_forward(x)
        got_feat = lore_engine.debug_tensor("feat")
        d = float((got - want).abs().max())
        print(f"fused DCN vs im2col + GEMM {shape}: max |d maps| = {d:.3e}, max |d feat| = {float((got_feat - want_feat).abs().max()):.3e}")
        assert torch.equal(got_feat, want_feat) and torch.equal(got, want)
    names = [r["kernel"] for r in _profile(lore_engine, x)]
    assert "dcn_fused_tcgen05" in names and "k_dcn_im2col" not in names
    assert "k_dcn_im2col" in [r["kernel"] for r in _profile(ref_eng, x)]
    ref_eng.close()


def _profile(eng, x):
    eng.profile_begin()
    eng.lore_detect_forward(x)
    return eng.profile_report()


def test_lore_detector_fp32x_meets_the_north_star_bound(post_engine):
    """precision="fp32x" (split-fp16 operand pairs through every conv, fp32 deformable sampling): head maps, DLA levels and the
    sparse ax / cr cell features within 1e-3 of the fp32 oracle (relative to each tensor's range, absolute for ranges < 1)."""
    TOL = 1e-3
    sd = synth.lore_dla34_state_dict(0)
    eng = Engine("lore_dla34", weights.pack_lore_dla34(sd, precise=True))
    g = np.load(os.path.join(GOLDEN, "lore_dla34_seed0.npz"))
    got = _unpack(eng.lore_detect_forward(torch.from_numpy(g["x"]).cuda()))
    for k in ("hm", "reg", "wh", "st"):
        want = g[k].transpose(0, 2, 3, 1)
        if k == "hm":
            want = 1.0 / (1.0 + np.exp(-want))
        err = float(np.abs(got[k] - want).max()) / max(1.0, float(np.abs(want).max()))
        print(f"fp32x {k}: rel max|err| {err:.3e}")
        assert err < TOL, k
    rng = np.random.default_rng(21)
    x = torch.from_numpy(rng.standard_normal((2, 3, 96, 160)).astype(np.float32))
    maps = eng.lore_detect_forward(x.cuda())
    base = lore_net_ref.dla34_base(sd, x)
    for lvl in range(0, 6):
        want = base[lvl].numpy()
        err = float(np.abs(eng.debug_tensor(f"level{lvl}").cpu().numpy() - want).max())
        print(f"fp32x level{lvl}: max|err| {err:.3e} (max|x| {float(np.abs(want).max()):.2f})")
        assert err < TOL * max(float(np.abs(want).max()), 1.0), f"level{lvl}"
    want = lore_net_ref.lore_dla34_features(sd, x).numpy()
    err = float(np.abs(eng.debug_tensor("feat").cpu().numpy() - want).max())
    print(f"fp32x feat: max|err| {err:.3e} (max|x| {float(np.abs(want).max()):.2f})")
    assert err < TOL * max(float(np.abs(want).max()), 1.0)
    # sparse ax / cr heads on planted cells
    planted = maps.clone()
    prng = np.random.default_rng(5)
    for i in range(2):
        for _ in range(20):
            planted[i, int(prng.integers(2, 22)), int(prng.integers(2, 38)), 0] = float(prng.uniform(0.5, 0.95))
    eye = np.tile(np.array([[1.0, 0, 0], [0, 1.0, 0]]), (2, 1, 1))
    dec = post_engine.lore_decode(planted, None, None, None, eye, wiz_rev=False, vis_thresh=0.3)
    feat, offsets = eng.lore_cell_features(dec, max_rows=128, check_overflow=True)
    out = lore_net_ref.lore_dla34_forward(sd, x, heads=("ax", "cr"))
    counts, offs = dec["counts"].cpu().numpy(), offsets.cpu().numpy()
    worst, scale = 0.0, 1.0
    for i in range(2):
        ax, cr = out["ax"][i].numpy().reshape(256, -1), out["cr"][i].numpy().reshape(256, -1)
        a_idx, c_idx = dec["ax_idx"].cpu().numpy()[i, : counts[i]], dec["cr_idx"].cpu().numpy()[i, : counts[i]]
        want = ax[:, a_idx].T + sum(cr[:, c_idx[:, k]].T for k in range(4))
        worst = max(worst, float(np.abs(feat.cpu().numpy()[offs[i]: offs[i + 1]] - want).max()))
        scale = max(scale, float(np.abs(want).max()))
    print(f"fp32x cell features: {counts.sum()} cells, max|err| {worst:.3e} (max|x| {scale:.2f})")
    assert counts.sum() >= 20 and worst < TOL * scale
    eng.close()


def test_lore_cell_features_vs_dense_heads(lore_engine, post_engine):
    """The sparse ax / cr evaluation equals gathering the oracle's dense head maps at the same points."""
    sd = synth.lore_dla34_state_dict(0)
    rng = np.random.default_rng(33)
    n, h, w = 2, 288, 320  # 72 x 80 maps: >= 5000 positions
    x = torch.from_numpy(rng.standard_normal((n, 3, h, w)).astype(np.float32))
    maps = lore_engine.lore_detect_forward(x.cuda())
    # random-weight heat maps have no sharp peaks above the gates: plant a few so that cells are selected
    planted = maps.clone()
    prng = np.random.default_rng(5)
    for i in range(n):
        for _ in range(40):
            y0, x0 = int(prng.integers(4, h // 4 - 4)), int(prng.integers(4, w // 4 - 4))
            planted[i, y0, x0, 0] = float(prng.uniform(0.5, 0.95))
    eye = np.tile(np.array([[1.0, 0, 0], [0, 1.0, 0]]), (n, 1, 1))
    dec = post_engine.lore_decode(planted, None, None, None, eye, wiz_rev=False, vis_thresh=0.3)
    feat, offsets = lore_engine.lore_cell_features(dec, max_rows=256, check_overflow=True)
    lore_engine.sync()
    out = lore_net_ref.lore_dla34_forward(sd, x, heads=("ax", "cr"))
    counts = dec["counts"].cpu().numpy()
    offs = offsets.cpu().numpy()
    assert counts.sum() >= 40 and offs[-1] == counts.sum()
    worst, scale = 0.0, 1.0
    for i in range(n):
        ax = out["ax"][i].numpy().reshape(256, -1)
        cr = out["cr"][i].numpy().reshape(256, -1)
        a_idx = dec["ax_idx"].cpu().numpy()[i, : counts[i]]
        c_idx = dec["cr_idx"].cpu().numpy()[i, : counts[i]]
        want = ax[:, a_idx].T + cr[:, c_idx[:, 0]].T + cr[:, c_idx[:, 1]].T + cr[:, c_idx[:, 2]].T + cr[:, c_idx[:, 3]].T
        got = feat.cpu().numpy()[offs[i]: offs[i + 1]]
        worst = max(worst, float(np.abs(got - want).max()))
        scale = max(scale, float(np.abs(want).max()))
    print(f"cell features: {counts.sum()} cells, max|err| {worst:.3e} (max|x| {scale:.2f})")
    assert worst < 2 * REL_TOL * scale  # sum of five head evaluations


def test_lore_processor_reference_golden(proc_engine):
    g = np.load(os.path.join(GOLDEN, "lore_processor_seed0.npz"))
    for n in (1, 7, 64, 200):
        feat = torch.zeros((256, 256), dtype=torch.float32)
        feat[:n] = torch.from_numpy(g[f"n{n}_feat"])
        offsets = torch.tensor([0, n], dtype=torch.int32)
        logic, stacked = proc_engine.lore_process_forward(feat.cuda(), offsets.cuda())
        proc_engine.sync()
        e1 = float(np.abs(logic.cpu().numpy()[:n] - g[f"n{n}_logic"]).max())
        e2 = float(np.abs(stacked.cpu().numpy()[:n] - g[f"n{n}_stacked"]).max())
        print(f"processor n={n}: max|err| logic {e1:.2e} stacked {e2:.2e}")
        assert e1 < PROC_TOL and e2 < PROC_TOL
        # structure tokens: identical wherever the reference is not within PROC_TOL of the .5 rounding boundary
        want = lore_decode_ref.round_logic(g[f"n{n}_stacked"])
        got = lore_decode_ref.round_logic(stacked.cpu().numpy()[:n])
        safe = np.abs((g[f"n{n}_stacked"] - np.floor(g[f"n{n}_stacked"])) - 0.5) > PROC_TOL
        np.testing.assert_array_equal(got[safe], want[safe])
        assert safe.mean() > 0.99


def test_lore_processor_segments(proc_engine):
    """Three images' cells packed in one row list: attention must stay inside each image's segment."""
    sd = synth.lore_processor_state_dict(0)
    rng = np.random.default_rng(8)
    sizes = [5, 0, 37, 90]
    rows = sum(sizes)
    feat = rng.standard_normal((rows, 256)).astype(np.float32)
    offsets = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int32)
    buf = torch.zeros((192, 256), dtype=torch.float32)
    buf[:rows] = torch.from_numpy(feat)
    logic, stacked = proc_engine.lore_process_forward(buf.cuda(), torch.from_numpy(offsets).cuda())
    proc_engine.sync()
    for i, sz in enumerate(sizes):
        if sz == 0:
            continue
        a, b = offsets[i], offsets[i + 1]
        wl, ws = lore_processor_ref.lore_processor_forward(sd, torch.from_numpy(feat[a:b]))
        assert float(np.abs(logic.cpu().numpy()[a:b] - wl.numpy()).max()) < PROC_TOL
        assert float(np.abs(stacked.cpu().numpy()[a:b] - ws.numpy()).max()) < PROC_TOL


def test_lore_ptn_configuration(post_engine):
    """task_type="ptn" (configuration_lore.py:101-116): 512 x 512 input, 3-layer transformers, the decode's integer position
    features looked up in the x / y embedding tables and added to the cell features (dv_lore_add_position_embeddings) -- against
    the processor oracle called with dets=, and end to end through OcrTableStructureTask."""
    from pdf_table_b200 import predictors

    psd = synth.lore_processor_state_dict(0, layers=3, stacking_layers=3)
    proc = Engine("lore_processor", weights.pack_lore_processor(psd))
    rng = np.random.default_rng(12)
    n_img, k, counts = 2, 40, np.array([23, 31], np.int32)
    feat = (rng.standard_normal((64, 256)) * 0.5).astype(np.float32)
    dets = rng.integers(0, 128, (n_img, k, 8)).astype(np.int32)
    offs = np.array([0, 23, 54], np.int32)
    dec = {"dets_feat": torch.from_numpy(dets).cuda(), "counts": torch.from_numpy(counts).cuda()}
    f_dev = torch.from_numpy(feat).cuda()
    proc.lore_add_position_embeddings(f_dev, dec, torch.from_numpy(offs).cuda())
    xe, ye = psd["x_position_embeddings.weight"], psd["y_position_embeddings.weight"]
    want = feat.copy()
    for i in range(n_img):
        d = dets[i, : counts[i]]
        want[offs[i]: offs[i + 1]] = (((feat[offs[i]: offs[i + 1]] + xe[d[:, 0]]) + ye[d[:, 1]]) + xe[d[:, 2]]) + ye[d[:, 5]]
    np.testing.assert_array_equal(f_dev.cpu().numpy(), want)  # fp32 adds in the reference's order: exact
    _, stacked = proc.lore_process_forward(f_dev, torch.from_numpy(offs).cuda())
    for i in range(n_img):
        _, ws = lore_processor_ref.lore_processor_forward(psd, torch.from_numpy(feat[offs[i]: offs[i + 1]]), layers=3, stacking_layers=3,
                                                          dets=torch.from_numpy(dets[i, : counts[i]].astype(np.int64)))
        err = float((stacked[offs[i]: offs[i + 1]].cpu() - ws).abs().max())
        print(f"ptn processor image {i}: max|err| {err:.2e}")
        assert err < PROC_TOL
    proc.close()
    # end to end
    sd = synth.lore_dla34_state_dict(0)
    sd["hm.2.bias"] = np.array([-0.3, -3.5], np.float32)
    task = predictors.OcrTableStructureTask(model="Lore", task_type="ptn", state_dict=(sd, psd))
    assert task.resolution == (512, 512) and task.vis_thresh == 0.35 and not task.wiz_rev
    res = task([synth.synthetic_page(7, 400, 600)])
    assert len(res) == 1 and res[0]["polygons"].shape[1] == 8 and res[0]["logi"].shape == (len(res[0]["polygons"]), 4)
    with pytest.raises(RuntimeError):
        predictors.OcrTableStructureTask(model="Lore", task_type="fin", state_dict=(sd, psd))
