"""tests/golden/crnn_seed0.npz: the reference CRNN module (crnn/modeling_crnn.py) run in the BUILD CONTAINER on seeded weights
and inputs, and the reference pre / post-processors of the recognition task on synthetic crops (see gen_golden.py)."""
from __future__ import annotations

import os

import numpy as np
import torch

from . import ref_import
from pdf_table_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
LABELS = 100  # small classifier for the fixture (the module's 7644-wide Linear is replaced; same forward code)


def case_input(n: int, w: int) -> np.ndarray:
    return np.random.default_rng(21).uniform(0.0, 1.0, (n, 3, 32, w)).astype(np.float32)


def main():
    ref_import.setup()
    from pdftable.model.crnn.modeling_crnn import CRNN

    model = CRNN().eval()
    model.cls = torch.nn.Linear(512, LABELS, bias=False)
    sd = synth.crnn_state_dict(0, LABELS)
    model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    missing = [k for k in model.state_dict() if k not in sd and not k.endswith("num_batches_tracked")]
    assert not missing, missing
    out = {}
    for n, w in ((2, 300), (1, 640), (3, 64)):
        x = torch.from_numpy(case_input(n, w))
        with torch.no_grad():
            y = model(x)
        out[f"logits_{n}x{w}"] = y.numpy()
        print(n, w, tuple(y.shape), float(y.abs().max()), float(y.std()))
    np.savez_compressed(os.path.join(GOLDEN, "crnn_seed0.npz"), **out)


if __name__ == "__main__":
    main()
