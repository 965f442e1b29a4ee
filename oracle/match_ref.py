"""Restatement of the reference's cell / text matching rule (TEST ORACLE, see oracle/__init__.py): find_top1_mach_box
(ocr_pdf/ocr_table_to_html_task.py:48-77) with box_in_other_box, distance and compute_iou_v2 (pdf_table/table_common.py:138-160,
435-441, 473-516), in Python floats exactly as the reference evaluates them.  Pinned against the reference's own functions by
tests/golden/match_seed0.npz (oracle/gen_golden_match.py executes their source from /root/reference)."""
from __future__ import annotations

from typing import List, Sequence


def box_in_other_box(box_1, box_2, diff=2):
    x1, y1, x2, y2 = box_1
    x3, y3, x4, y4 = box_2
    min_y_1, max_y_1 = min(y1, y2), max(y1, y2)
    min_y_2, max_y_2 = min(y3, y4), max(y3, y4)
    return bool(x3 >= x1 - diff and x4 <= x2 + diff and min_y_1 - diff <= min_y_2 <= max_y_2 <= max_y_1 + diff)


def distance(box_1, box_2):
    x1, y1, x2, y2 = box_1
    x3, y3, x4, y4 = box_2
    dis = abs(x3 - x1) + abs(y3 - y1) + abs(x4 - x2) + abs(y4 - y2)
    return dis + min(abs(x3 - x1) + abs(y3 - y1), abs(x4 - x2) + abs(y4 - y2))


def compute_iou_v2(a, b):
    x1, y1, x2, y2 = max(a[0], b[0]), max(a[1], b[1]), min(a[2], b[2]), min(a[3], b[3])
    dx, dy = max(x2 - x1, 0), max(y2 - y1, 0)
    inter = dx * dy
    return inter / (abs((a[2] - a[0]) * (a[3] - a[1])) + abs((b[2] - b[0]) * (b[3] - b[1])) - inter + 1e-6)


def find_top1(text_box: Sequence[float], cells: Sequence[Sequence[float]]) -> int:
    distances = []
    for index, cell in enumerate(cells):
        if box_in_other_box(cell, text_box):
            return index
        distances.append((distance(text_box, cell), 1.0 - compute_iou_v2(text_box, cell)))
    best = sorted(distances, key=lambda item: (item[1], item[0]))[0]
    return distances.index(best)


def match(text_boxes, cells) -> List[int]:
    return [find_top1([float(v) for v in t], [[float(v) for v in c] for c in cells]) for t in text_boxes]
