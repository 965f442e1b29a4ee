"""CTC greedy decode kernel vs the reference-generated golden strings and the numpy oracle (bit-exact)."""
import os

import numpy as np
import pytest
import torch

from oracle import ctc_ref

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _run(eng, preds):
    ids, ln, conf = eng.ctc_greedy(torch.from_numpy(preds).cuda())
    eng.sync()
    ids, ln, conf = ids.cpu().numpy(), ln.cpu().numpy(), conf.cpu().numpy()
    return [ids[b, : ln[b]] for b in range(len(ln))], ids, ln, conf


def test_ctc_golden_strings(post_engine):
    g = np.load(os.path.join(GOLDEN, "ctc_decode.npz"))
    character = list(g["character"])
    for case in ("known", "rand_T40", "rand_T7", "rand_T160", "rand_T300", "edge"):
        rows, ids, ln, conf = _run(post_engine, g[f"{case}.preds"])
        texts = ["".join(character[i] for i in r) for r in rows]
        assert texts == list(g[f"{case}.text"]), case
        # confidence is bit-exact float32 (numpy pairwise mean)
        np.testing.assert_array_equal(conf.astype(np.float64), g[f"{case}.conf"], err_msg=case)
        for b in range(len(ln)):
            assert (ids[b, ln[b]:] == -1).all()


@pytest.mark.parametrize("B,T,C", [(64, 40, 97), (16, 40, 6625), (3, 1, 5), (7, 321, 33), (2, 1024, 131)])
def test_ctc_vs_oracle_seeded(post_engine, B, T, C):
    rng = np.random.default_rng(B * 1000 + T + C)
    logits = rng.standard_normal((B, T, C)).astype(np.float32) * 3
    runs = rng.integers(0, C, size=(B, T))
    for t in range(1, T):
        same = rng.random(B) < 0.4
        runs[same, t] = runs[same, t - 1]
        runs[rng.random(B) < 0.2, t] = 0
    logits[np.arange(B)[:, None], np.arange(T)[None, :], runs] += 6
    e = np.exp(logits - logits.max(-1, keepdims=True))
    preds = (e / e.sum(-1, keepdims=True)).astype(np.float32)
    want_ids, want_conf = ctc_ref.ctc_greedy_ids(preds)
    rows, ids, ln, conf = _run(post_engine, preds)
    for b in range(B):
        np.testing.assert_array_equal(rows[b], want_ids[b])
    np.testing.assert_array_equal(conf, want_conf)


def test_ctc_empty_batch(post_engine):
    ids, ln, conf = post_engine.ctc_greedy(torch.zeros((0, 8, 11), device="cuda"))
    assert ids.shape == (0, 8) and ln.numel() == 0
