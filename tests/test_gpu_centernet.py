"""CenterNet table structure on the engine: DLA-34 + plain IDA-up network, vertex-grouping decode, task mirror."""
import os

import numpy as np
import pytest
import torch

from oracle import centernet_ref
from pdf_table_b200 import predictors, synth, weights
from pdf_table_b200.engine import Engine

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DECODE_CASES = [("c0", 20, 128, 128, (600, 800)), ("c1", 21, 128, 160, (1024, 1280)), ("c2", 22, 256, 256, (1500, 1100))]
REL_TOL = 4e-3  # max|err| relative to max(1, max|oracle|) per head map (fp16 operands, ~60 conv layers)


@pytest.fixture(scope="module")
def cn_engine():
    eng = Engine("centernet_dla34", weights.pack_centernet_dla34(synth.centernet_dla34_state_dict(0)))
    yield eng
    eng.close()


def test_centernet_network_reference_golden(cn_engine):
    g = np.load(os.path.join(GOLDEN, "centernet_dla34_seed0.npz"))
    maps = cn_engine.lore_detect_forward(torch.from_numpy(g["x"]).cuda()).cpu().numpy()
    cn_engine.sync()
    got = {"hm": maps[..., 0:2], "reg": maps[..., 2:4], "c2v": maps[..., 4:12], "v2c": maps[..., 12:20]}
    for k in ("hm", "reg", "c2v", "v2c"):
        want = g[k].transpose(0, 2, 3, 1)
        if k == "hm":
            want = 1.0 / (1.0 + np.exp(-want))
        rel = float(np.abs(got[k] - want).max()) / max(1.0, float(np.abs(want).max()))
        print(f"centernet {k}: rel max|err| {rel:.2e} (max|x| {float(np.abs(want).max()):.2f})")
        assert rel < REL_TOL, k


def test_centernet_decode_reference_golden(post_engine):
    g = np.load(os.path.join(GOLDEN, "centernet_decode.npz"))
    for name, idx, h, w, (sh, sw) in DECODE_CASES:
        m = synth.lore_planted_maps(idx, h, w, with_feat=False)
        inv = predictors.lore_affine(np.array([sw / 2.0, sh / 2.0], np.float32), max(sh, sw) * 1.0, w, h, inv=True)
        polygons, counts = post_engine.centernet_decode(*[torch.from_numpy(m[k])[None].cuda() for k in ("hm", "reg", "wh", "st")], inv[None])
        post_engine.sync()
        n = int(counts.cpu()[0])
        assert n == len(g[name]) and n > 0, name
        np.testing.assert_array_equal(polygons.cpu().numpy()[0, :n], g[name], err_msg=name)


def test_centernet_decode_batch_vs_oracle(post_engine):
    cases = [(30, (700, 900)), (31, (1024, 1024)), (32, (400, 1300))]
    h = w = 128
    packed = np.zeros((len(cases), h, w, 24), np.float32)
    want, inv = [], []
    for i, (idx, (sh, sw)) in enumerate(cases):
        m = synth.lore_planted_maps(idx, h, w, with_feat=False)
        packed[i, :, :, 0:2], packed[i, :, :, 2:4] = m["hm"].transpose(1, 2, 0), m["reg"].transpose(1, 2, 0)
        packed[i, :, :, 4:12], packed[i, :, :, 12:20] = m["wh"].transpose(1, 2, 0), m["st"].transpose(1, 2, 0)
        c, s = np.array([sw / 2.0, sh / 2.0], np.float32), max(sh, sw) * 1.0
        want.append(centernet_ref.centernet_decode(m["hm"], m["reg"], m["wh"], m["st"], c, s, h, w))
        inv.append(predictors.lore_affine(c, s, w, h, inv=True))
    polygons, counts = post_engine.centernet_decode(torch.from_numpy(packed).cuda(), None, None, None, np.stack(inv))
    post_engine.sync()
    for i in range(len(cases)):
        n = int(counts.cpu()[i])
        assert n == len(want[i])
        np.testing.assert_array_equal(polygons.cpu().numpy()[i, :n], want[i])
    z2, z8 = torch.zeros(1, 2, 80, 80).cuda(), torch.zeros(1, 8, 80, 80).cuda()
    _, counts = post_engine.centernet_decode(z2, z2, z8, z8, np.array([[[1.0, 0, 0], [0, 1.0, 0]]]))
    assert int(counts.cpu()[0]) == 0


def test_table_structure_task_centernet():
    sd = synth.centernet_dla34_state_dict(0)
    sd["hm.2.bias"] = np.array([0.5, -1.0], np.float32)
    task = predictors.OcrTableStructureTask(model="CenterNet", state_dict=sd)
    pages = [synth.synthetic_page(7, 700, 900), synth.synthetic_page(8, 1024, 768)]
    res = task(pages)
    assert len(res) == 2 and all(r["polygons"].ndim == 2 and r["polygons"].shape[1] == 8 for r in res)
    # the decode of the engine's own maps equals the oracle decode bit for bit
    pre = task._preprocess(pages)
    maps = task.predictor.lore_detect_forward_u8(pre["images"]).cpu().numpy()  # the warp ran on the device (bit-exact vs cv2)
    for i, (c, s) in enumerate(pre["cs"]):
        want = centernet_ref.centernet_decode(maps[i, :, :, 0:2].transpose(2, 0, 1), maps[i, :, :, 2:4].transpose(2, 0, 1),
                                              maps[i, :, :, 4:12].transpose(2, 0, 1), maps[i, :, :, 12:20].transpose(2, 0, 1), c, s, 256, 256)
        np.testing.assert_array_equal(res[i]["polygons"], want)
    print("centernet task: cells per page", [len(r["polygons"]) for r in res])
