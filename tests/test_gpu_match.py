"""SURVEY.md 8(f)-4 on the GPU: cell / text matching (dv_match_cells) against the indices of the reference's own functions
(tests/golden/match_seed0.npz) and against the oracle restatement on larger random tables.  Index work: exact."""
import os

import numpy as np
import pytest
import torch

from oracle import match_ref
from pdf_table_b200 import system
from pdf_table_b200.engine import Engine

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "match_seed0.npz")


def test_match_cells_equals_reference_golden(post_engine):
    g = np.load(GOLDEN)
    for i in range(6):
        texts, cells = g[f"texts{i}"], g[f"cells{i}"]
        got = post_engine.match_cells(torch.from_numpy(texts).cuda(), torch.from_numpy(cells).cuda()).cpu().numpy()
        np.testing.assert_array_equal(got, g[f"top1_{i}"])
        matched = system.match_table_cell_and_text_cell(post_engine, cells, texts)
        want = {}
        for k, c in enumerate(g[f"top1_{i}"].tolist()):
            want.setdefault(c, []).append(k)
        assert matched == want and list(matched) == list(want)


def test_match_cells_vs_oracle_on_large_tables(post_engine):
    rng = np.random.default_rng(4)
    for n_cells, n_text in ((1, 5), (33, 200), (700, 1500)):
        x0, y0 = rng.uniform(0, 2000, n_cells), rng.uniform(0, 2000, n_cells)
        cells = np.stack([x0, y0, x0 + rng.uniform(5, 300, n_cells), y0 + rng.uniform(5, 80, n_cells)], 1)
        cells[::7] = np.round(cells[::7])
        tx, ty = rng.uniform(-20, 2100, n_text), rng.uniform(-20, 2100, n_text)
        texts = np.stack([tx, ty, tx + rng.uniform(3, 200, n_text), ty + rng.uniform(3, 40, n_text)], 1)
        texts[: n_text // 3] = cells[rng.integers(0, n_cells, n_text // 3)] + rng.uniform(-2, 2, (n_text // 3, 4))  # around the diff margin
        got = post_engine.match_cells(torch.from_numpy(texts).cuda(), torch.from_numpy(cells).cuda()).cpu().numpy().tolist()
        assert got == match_ref.match(texts, cells)
    assert system.match_table_cell_and_text_cell(post_engine, [[0, 0, 1, 1]], []) == {}
