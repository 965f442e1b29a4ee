"""BASELINE configs[1] at FULL size on the GPU (32 pages 960 x 960, 1280 text-line crops) through properties that need no
oracle run of that size: determinism, independence of an item's result from the batch it is processed in, equal inputs ->
equal outputs inside one batch, permutation equivariance.  The small-size parity tests pin the values; these pin that nothing
changes when the batch grows to the size the benchmark is quoted on."""
import numpy as np
import pytest
import torch

from pdf_table_b200 import synth, weights
from pdf_table_b200.engine import Engine

pytestmark = pytest.mark.gpu
N_PAGES, PAGE, CROPS_PER_PAGE = 32, 960, 40
MEAN, STD = (0.485, 0.456, 0.406), (0.229, 0.224, 0.225)
PROB_TOL = 1e-2  # the probability-map tolerance of the parity tests (fp16 operands); measured differences here should be 0


def _tiled(make, n, distinct):
    items = [make(i) for i in range(distinct)]
    return np.stack([items[i % distinct] for i in range(n)])


def test_dbnet_full_batch_is_deterministic_and_batch_independent():
    det = Engine("dbnet_r18", weights.pack_dbnet_r18(synth.dbnet_r18_state_dict(0)))
    pages = torch.from_numpy(_tiled(lambda i: synth.synthetic_page(400 + i, PAGE, PAGE), N_PAGES, 4)).cuda()
    full = det.dbnet_forward_u8(pages, MEAN, STD, 1.0 / 255.0, True).clone()
    again = det.dbnet_forward_u8(pages, MEAN, STD, 1.0 / 255.0, True)
    assert full.shape == (N_PAGES, 1, PAGE, PAGE) and torch.isfinite(full).all()
    assert torch.equal(full, again)
    assert float((full[1] - full[5]).abs().max()) <= PROB_TOL  # pages 1 and 5 are the same image
    for k in (0, 13, 31):
        one = det.dbnet_forward_u8(pages[k:k + 1].contiguous(), MEAN, STD, 1.0 / 255.0, True)
        err = float((one[0] - full[k]).abs().max())
        print(f"page {k}: alone vs in the batch of {N_PAGES}: max|dprob| = {err:.2e}")
        assert err <= PROB_TOL
    det.close()


def test_db_boxes_full_batch_equals_per_page():
    post = Engine("post")
    maps = torch.from_numpy(_tiled(lambda i: synth.synthetic_prob_map(500 + i, PAGE, PAGE, CROPS_PER_PAGE), N_PAGES, 4)[:, None]).cuda()
    src = [(PAGE, PAGE)] * N_PAGES
    boxes, counts = post.db_boxes(maps, src)
    boxes, counts = boxes.cpu().numpy(), counts.cpu().numpy()
    assert (counts > 10).all()
    for k in range(4, N_PAGES):  # equal maps -> equal boxes wherever they sit in the batch
        assert counts[k] == counts[k % 4] and np.array_equal(boxes[k, :counts[k]], boxes[k % 4, :counts[k]])
    for k in (0, 17, 31):
        b1, c1 = post.db_boxes(maps[k:k + 1].contiguous(), [(PAGE, PAGE)])
        assert int(c1[0]) == counts[k] and np.array_equal(b1.cpu().numpy()[0, :counts[k]], boxes[k, :counts[k]])
    post.close()


def test_ctc_greedy_full_batch_is_permutation_equivariant():
    post = Engine("post")
    rng = np.random.default_rng(61)
    n = N_PAGES * CROPS_PER_PAGE
    logits = rng.standard_normal((n, 40, 97)).astype(np.float32) * 3
    e = np.exp(logits - logits.max(-1, keepdims=True))
    probs = torch.from_numpy((e / e.sum(-1, keepdims=True)).astype(np.float32)).cuda()
    perm = torch.from_numpy(rng.permutation(n)).cuda()
    ids, ln, conf = post.ctc_greedy(probs)
    ids_p, ln_p, conf_p = post.ctc_greedy(probs[perm].contiguous())
    assert torch.equal(ids[perm], ids_p) and torch.equal(ln[perm], ln_p) and torch.equal(conf[perm], conf_p)
    assert int(ln.min()) >= 0 and int(ln.max()) <= 40
    post.close()


def test_recogniser_full_batch_is_batch_independent():
    rec = Engine("convnext_vit", weights.pack_convnext_vit(synth.convnext_vit_state_dict(0)))
    n = N_PAGES * CROPS_PER_PAGE
    crops = torch.from_numpy(_tiled(lambda i: synth.synthetic_text_crop(600 + i, 32, 320), n, 64)).cuda()
    full = rec.convnextvit_forward_u8(crops).clone()
    assert full.shape == (n, 201)
    same = float((full[:64] == full[640:704]).float().mean())  # crops i and i + 640 are the same image
    sub = rec.convnextvit_forward_u8(crops[100:140].contiguous())
    indep = float((sub == full[100:140]).float().mean())
    print(f"equal crops -> equal ids: {same:.5f}; sub-batch vs batch of {n}: {indep:.5f}")
    # identical arithmetic per token is expected (1.0); the bound only allows arg-max flips at exact near-ties
    assert same >= 0.999 and indep >= 0.999
    rec.close()
