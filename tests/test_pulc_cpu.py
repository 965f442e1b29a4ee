"""SURVEY.md 8(f)-3 without a GPU: the PULC classifier's oracle restatement, the graph-program lowering and the host pre / post
steps of ClsImagePulcTask against goldens made from the reference's own PPLCNet module and image processors
(oracle/gen_golden_pulc.py -> tests/golden/pulc_seed0.npz)."""
import json
import os

import numpy as np
import torch

from oracle import graph_interp, pplcnet_ref
from oracle.gen_golden_pulc import CASES, case_input, case_logits
from pdf_table_b200 import pplcnet_graph as G
from pdf_table_b200 import predictors, synth

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pulc_seed0.npz")


def test_oracle_and_lowering_reproduce_the_reference_module():
    g = np.load(GOLDEN)
    for task, n, h, w in CASES:
        sd = synth.pplcnet_cls_state_dict(0, G.TASK_CLASSES[task])
        x = torch.from_numpy(case_input(n, h, w))
        want = pplcnet_ref.pplcnet_forward(sd, x, G.TASK_STRIDES[task])
        np.testing.assert_allclose(want.numpy(), g[task + ".logits"], atol=2e-6, rtol=0)  # oracle == reference PPLCNet
        blob, meta = G.build_pplcnet(sd, G.TASK_STRIDES[task])
        assert meta["class_num"] == G.TASK_CLASSES[task] and int(blob["graph.meta"][5]) == 2
        _, heads = graph_interp.run_program(blob, x, fp16_activations=True)
        assert float((heads["logits"][:, 0] - want).abs().max()) < 5e-3  # fp16 weights + activation buffers


class _HostOnly(predictors.ClsImagePulcTask):
    """The host halves of the task (no engine): enough for the pre / post-processing goldens."""

    def __init__(self, task_type):
        self.task_type = task_type


def test_host_pre_and_post_processing_equal_the_reference_processors():
    g = np.load(GOLDEN)
    want_post = json.loads(bytes(g["post_json"]).decode())
    page = synth.synthetic_page(3, 120, 300)
    for task, _, _, _ in CASES:
        t = _HostOnly(task)
        pv = t._preprocess(page)["pixel_values"]
        np.testing.assert_array_equal(pv[:, :, ::4, ::4], g[task + ".pixel_values"])
        assert float(pv.astype(np.float64).sum()) == float(g[task + ".pixel_sum"].ravel()[0])
        got = t._postprocess({"logits": case_logits(G.TASK_CLASSES[task])})
        assert got == want_post[task]
    assert isinstance(_HostOnly("textline_orientation")._postprocess({"logits": case_logits(2)[:1]}), dict)  # one input -> bare dict
