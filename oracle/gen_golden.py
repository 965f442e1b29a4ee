"""Generate tests/golden/*.npz by running the REFERENCE's own modules (build container only).

    python -m oracle.gen_golden            # writes tests/golden/

Fixtures are small (KBs) and committed together with this script; the GPU box only reads them.
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import torch

from . import ref_import
from pdf_table_b200 import synth

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def gen_dbnet():
    """Reference DBModel (db_net/dbnet.py:715) with the seeded synthetic state_dict on a 1x3x64x96 input."""
    DBModel = ref_import.dbmodel()
    sd = synth.dbnet_r18_state_dict(0)
    model = DBModel().eval()
    missing, unexpected = model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()}, strict=False)
    assert not unexpected, unexpected
    assert all(k.startswith("decoder.thresh") or k.endswith("num_batches_tracked") for k in missing), missing
    rng = np.random.default_rng(7)
    x = rng.standard_normal((1, 3, 64, 96)).astype(np.float32)
    with torch.no_grad():
        y = model(torch.from_numpy(x)).numpy()
    np.savez_compressed(os.path.join(GOLDEN, "dbnet_r18_seed0.npz"), x=x, prob=y.astype(np.float32))
    print("dbnet_r18_seed0", y.shape, float(y.min()), float(y.max()), float(y.mean()))


def gen_ctc():
    """Reference CTCLabelDecode (ocr_rec_pp/rec_postprocess.py:167-195) on seeded softmax tensors."""
    CTC = ref_import.ctc_label_decode()
    chars = [chr(ord("0") + i) for i in range(10)] + [chr(ord("a") + i) for i in range(26)] + list("ABCDEFGHIJ")
    with tempfile.NamedTemporaryFile("w", suffix=".txt", delete=False, encoding="utf-8") as f:
        f.write("\n".join(chars) + "\n")
        path = f.name
    dec = CTC(character_dict_path=path, use_space_char=True)
    os.unlink(path)
    C = len(dec.character)
    out = {"character": np.array(dec.character)}
    rng = np.random.default_rng(11)
    cases = {}
    # (i) SURVEY.md 8c known-answer case: ids [1,1,0,2,2,2,0,0,3,C-1,0...] at p=0.9 -> '012 '
    T = 16
    ids = np.zeros(T, np.int64)
    ids[:10] = [1, 1, 0, 2, 2, 2, 0, 0, 3, C - 1]
    p = np.full((1, T, C), 0.1 / (C - 1), np.float32)
    p[0, np.arange(T), ids] = 0.9
    cases["known"] = p
    # (ii) random peaked softmax, with repeats and blanks, several T
    for name, (B, T2, temp) in {"rand_T40": (8, 40, 6.0), "rand_T7": (5, 7, 4.0), "rand_T160": (3, 160, 8.0),
                               "rand_T300": (2, 300, 8.0)}.items():
        logits = rng.standard_normal((B, T2, C)).astype(np.float32) * temp
        # force runs and blanks
        runs = rng.integers(0, C, size=(B, T2))
        for b in range(B):
            for t in range(1, T2):
                r = rng.random()
                if r < 0.35:
                    runs[b, t] = runs[b, t - 1]
                elif r < 0.6:
                    runs[b, t] = 0
        logits[np.arange(B)[:, None], np.arange(T2)[None, :], runs] += 2 * temp
        e = np.exp(logits - logits.max(-1, keepdims=True))
        cases[name] = (e / e.sum(-1, keepdims=True)).astype(np.float32)
    # (iii) all-blank rows and exact ties (argmax must take the first maximum)
    p = np.zeros((2, 5, C), np.float32)
    p[0, :, 0] = 1.0
    p[1, :, :] = 1.0 / C
    p[1, 2, 5] = p[1, 2, 9] = 0.3
    cases["edge"] = p
    for name, p in cases.items():
        res = dec(p)
        out[f"{name}.preds"] = p
        out[f"{name}.text"] = np.array([r[0] for r in res])
        out[f"{name}.conf"] = np.array([r[1] for r in res], np.float64)
        print("ctc", name, [r[0][:24] for r in res][:3])
    np.savez_compressed(os.path.join(GOLDEN, "ctc_decode.npz"), **out)


def main(which=None):
    os.makedirs(GOLDEN, exist_ok=True)
    gens = {"dbnet": gen_dbnet, "ctc": gen_ctc}
    try:
        from . import gen_golden_more
        gens.update(gen_golden_more.GENERATORS)
    except ImportError:
        pass
    for name, fn in gens.items():
        if which and name not in which:
            continue
        fn()


if __name__ == "__main__":
    main(sys.argv[1:] or None)
