"""tests/golden/match_seed0.npz: the reference's OWN matching functions executed in the BUILD CONTAINER.  The modules that hold
them import pdfminer (absent here), so their source is cut out of the files with `ast` and executed as is: box_in_other_box,
distance, compute_iou_v2 (pdf_table/table_common.py) and the method find_top1_mach_box (ocr_pdf/ocr_table_to_html_task.py)."""
from __future__ import annotations

import ast
import os

import numpy as np

from . import ref_import

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def _source_of(path: str, names) -> str:
    src = open(path).read()
    tree = ast.parse(src)
    out = []
    for node in ast.walk(tree):
        if isinstance(node, ast.FunctionDef) and node.name in names:
            out.append(ast.get_source_segment(src, node))
    assert len(out) == len(names), (names, len(out))
    return "\n\n".join(out)


def cases(seed: int = 0):
    """Seeded tables: a grid of cells (jittered, some spanning), text boxes inside cells, straddling borders and outside the table;
    integer and fractional coordinates."""
    rng = np.random.default_rng(seed)
    out = []
    for k in range(6):
        rows, cols = int(rng.integers(2, 9)), int(rng.integers(2, 8))
        xs = np.cumsum(rng.integers(30, 120, cols + 1)).astype(np.float64)
        ys = np.cumsum(rng.integers(15, 60, rows + 1)).astype(np.float64)
        cells = []
        for r in range(rows):
            for c in range(cols):
                j = rng.uniform(-1.5, 1.5, 4) if k % 2 else np.zeros(4)
                cells.append([xs[c] + j[0], ys[r] + j[1], xs[c + 1] + j[2], ys[r + 1] + j[3]])
        cells = np.array(cells)
        if k >= 3:
            cells = cells[rng.permutation(len(cells))]
        texts = []
        for _ in range(int(rng.integers(10, 60))):
            mode = rng.integers(0, 4)
            c = cells[rng.integers(0, len(cells))]
            if mode == 0:  # well inside a cell
                w, h = (c[2] - c[0]) * rng.uniform(0.2, 0.8), (c[3] - c[1]) * rng.uniform(0.3, 0.8)
                x0, y0 = c[0] + rng.uniform(0, (c[2] - c[0]) - w), c[1] + rng.uniform(0, (c[3] - c[1]) - h)
            elif mode == 1:  # straddling the right / bottom border
                w, h = (c[2] - c[0]) * rng.uniform(0.5, 1.4), (c[3] - c[1]) * rng.uniform(0.5, 1.3)
                x0, y0 = c[0] + rng.uniform(0, 20), c[1] + rng.uniform(0, 10)
            elif mode == 2:  # touching the diff = 2 margin
                w, h = c[2] - c[0] + rng.choice([3.0, 4.0, 4.5]), c[3] - c[1]
                x0, y0 = c[0] - 2.0, c[1]
            else:  # anywhere, possibly outside the table
                w, h = rng.uniform(10, 200), rng.uniform(8, 40)
                x0, y0 = rng.uniform(-50, xs[-1] + 50), rng.uniform(-30, ys[-1] + 30)
            t = [x0, y0, x0 + w, y0 + h]
            texts.append([float(round(v)) for v in t] if k % 3 == 0 else t)
        out.append((np.array(texts, np.float64), cells.astype(np.float64)))
    return out


def main():
    assert ref_import.available()
    root = os.path.join(ref_import.R, "model")
    ns = {}
    exec(_source_of(os.path.join(root, "pdf_table", "table_common.py"), ["box_in_other_box", "distance", "compute_iou_v2"]), ns)
    exec(_source_of(os.path.join(root, "ocr_pdf", "ocr_table_to_html_task.py"), ["find_top1_mach_box"]), ns)

    class Cell:
        def __init__(self, b):
            self.b = [float(v) for v in b]

        def to_bbox(self):
            return self.b

    out = {}
    for i, (texts, cells) in enumerate(cases()):
        cl = [Cell(c) for c in cells]
        top1 = [ns["find_top1_mach_box"](None, text_box=[float(v) for v in t], table_bboxs=cl) for t in texts]
        out[f"texts{i}"], out[f"cells{i}"], out[f"top1_{i}"] = texts, cells, np.array(top1, np.int32)
        print(i, len(texts), len(cells), top1[:8])
    np.savez_compressed(os.path.join(GOLDEN, "match_seed0.npz"), **out)


if __name__ == "__main__":
    main()
