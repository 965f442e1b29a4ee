"""SURVEY.md 8(f)-3 on the GPU: the PULC PP-LCNet classifiers through dv_cls_forward against the reference PPLCNet module's
golden logits (tests/golden/pulc_seed0.npz) and the fp32 oracle, and ClsImagePulcTask end to end."""
import os

import numpy as np
import pytest
import torch

from oracle import pplcnet_ref
from oracle.gen_golden_pulc import CASES, case_input
from pdf_table_b200 import pplcnet_graph as G
from pdf_table_b200 import predictors, synth
from pdf_table_b200.engine import Engine

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pulc_seed0.npz")
LOGIT_TOL = 1e-2  # fp16 operands / activations; logits are O(1)


def test_pplcnet_cls_vs_reference_golden():
    g = np.load(GOLDEN)
    for task, n, h, w in CASES:
        sd = synth.pplcnet_cls_state_dict(0, G.TASK_CLASSES[task])
        eng = Engine("pplcnet_cls", G.pack_pplcnet(sd, G.TASK_STRIDES[task]))
        logits, probs = eng.cls_forward(torch.from_numpy(case_input(n, h, w)).cuda(), return_probs=True)
        err = float(np.abs(logits.cpu().numpy() - g[task + ".logits"]).max())
        print(f"pulc {task}: max |dlogit| = {err:.3e}")
        assert err <= LOGIT_TOL
        np.testing.assert_allclose(probs.cpu().numpy(), torch.softmax(logits.cpu(), -1).numpy(), atol=1e-6, rtol=0)
        eng.close()


def test_cls_image_pulc_task_end_to_end():
    pages = [synth.synthetic_page(3, 120, 300), synth.synthetic_page(4, 200, 260)]
    for task in ("textline_orientation", "table_attribute"):
        sd = synth.pplcnet_cls_state_dict(0, G.TASK_CLASSES[task])
        t = predictors.ClsImagePulcTask(task_type=task, state_dict=sd)
        res = t(pages)
        assert isinstance(res, list) and len(res) == 2
        pv = t._preprocess(pages)["pixel_values"]
        want = pplcnet_ref.pplcnet_forward(sd, torch.from_numpy(pv), G.TASK_STRIDES[task]).numpy()
        ref = t._postprocess({"logits": want})
        for a, b, lg in zip(res, ref, want):
            if task == "table_attribute":
                safe = np.abs(lg - 0.5) > LOGIT_TOL
                assert [x for x, s in zip(a["output"], safe) if s] == [x for x, s in zip(b["output"], safe) if s]
            else:
                assert a["class_ids"] == b["class_ids"] and a["label_names"] == b["label_names"]
                assert np.abs(np.array(a["scores"]) - np.array(b["scores"])).max() <= LOGIT_TOL
        assert isinstance(t(pages[0]), dict)  # a single input comes back as the bare dict, as the reference does
    with pytest.raises(RuntimeError):
        predictors.ClsImagePulcTask(task_type="nope", state_dict={})
