"""CPU restatement of the reference DB post-process "threshold + box seed" (TEST ORACLE, see oracle/__init__.py).

Follows model/db_pp/processor_ocr_db_pp.py: DBPostProcess.__call__ :291-311, boxes_from_bitmap :174-219,
get_mini_boxes :230-251, box_score_fast :253-268, unclip :221-228, PPOcrDetectionPostProcessor.__call__ :330-342,
order_points_clockwise :344-366, clip_det_res :368-372, filter_tag_det_res :374-386.

OpenCV (cv2, present in this image and on the GPU box) is called exactly where the reference calls it.  Two
third-party dependencies of the reference are ABSENT from the image and are restated here from their published
algorithms (parity for these two is therefore unpinned by a run of the real libraries, stated in DESIGN.md):
  * pyclipper 1.3.x (Angus Johnson's Clipper 6.4.2): PyclipperOffset().AddPath(box, JT_ROUND, ET_CLOSEDPOLYGON)
    .Execute(d) -- integer truncation of the path (pyclipper casts to cInt), ClipperOffset::DoOffset / OffsetPoint /
    DoRound with arc_tolerance 0.25, miter_limit 2.  The final Clipper union of the single convex offset polygon
    only removes duplicate / collinear vertices, which cv2.minAreaRect is invariant to, so it is not restated.
  * shapely Polygon(box).area / .length: shoelace area and perimeter in float64.
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import cv2
import numpy as np


# ------------------------------------------------------------------------------------------- Clipper restatement
def _cround(v: float) -> int:
    """clipper.cpp Round(): static_cast<cInt>(val -/+ 0.5) (truncation toward zero)."""
    return int(v - 0.5) if v < 0 else int(v + 0.5)


def _area(path: Sequence[Tuple[int, int]]) -> float:
    """clipper.cpp Area(): a += (poly[j].X + poly[i].X) * (poly[j].Y - poly[i].Y), j = previous; return -a/2."""
    a = 0.0
    j = len(path) - 1
    for i in range(len(path)):
        a += (float(path[j][0]) + path[i][0]) * (float(path[j][1]) - path[i][1])
        j = i
    return -a * 0.5


def clipper_offset_round(path: Sequence[Tuple[int, int]], delta: float, arc_tolerance: float = 0.25) -> List[Tuple[int, int]]:
    """ClipperOffset (jtRound, etClosedPolygon) of one integer path by `delta`; returns the raw m_destPoly."""
    # AddPath: strip trailing duplicates of the first point and consecutive duplicates
    pts = [tuple(int(c) for c in p) for p in path]
    hi = len(pts) - 1
    while hi > 0 and pts[0] == pts[hi]:
        hi -= 1
    src = [pts[0]]
    for i in range(1, hi + 1):
        if src[-1] != pts[i]:
            src.append(pts[i])
    if len(src) < 3:
        return []
    # FixOrientations: a single closed polygon with negative area is reversed
    if not (_area(src) >= 0):
        src = src[::-1]
    if abs(delta) < 1e-20:
        return list(src)
    if arc_tolerance <= 0.0:
        y = 0.25
    elif arc_tolerance > abs(delta) * 0.25:
        y = abs(delta) * 0.25
    else:
        y = arc_tolerance
    steps = math.pi / math.acos(1 - y / abs(delta))
    if steps > abs(delta) * math.pi:
        steps = abs(delta) * math.pi
    m_sin = math.sin(2 * math.pi / steps)
    m_cos = math.cos(2 * math.pi / steps)
    steps_per_rad = steps / (2 * math.pi)
    if delta < 0:
        m_sin = -m_sin
    n = len(src)

    def unit_normal(p1, p2):
        if p1 == p2:
            return (0.0, 0.0)
        dx, dy = float(p2[0] - p1[0]), float(p2[1] - p1[1])
        f = 1.0 / math.sqrt(dx * dx + dy * dy)
        dx *= f
        dy *= f
        return (dy, -dx)

    normals = [unit_normal(src[j], src[(j + 1) % n]) for j in range(n)]
    dest: List[Tuple[int, int]] = []
    k = n - 1
    for j in range(n):
        sin_a = normals[k][0] * normals[j][1] - normals[j][0] * normals[k][1]
        done = False
        if abs(sin_a * delta) < 1.0:
            cos_a = normals[k][0] * normals[j][0] + normals[j][1] * normals[k][1]
            if cos_a > 0:
                dest.append((_cround(src[j][0] + normals[k][0] * delta), _cround(src[j][1] + normals[k][1] * delta)))
                done = True
        elif sin_a > 1.0:
            sin_a = 1.0
        elif sin_a < -1.0:
            sin_a = -1.0
        if not done:
            if sin_a * delta < 0:
                dest.append((_cround(src[j][0] + normals[k][0] * delta), _cround(src[j][1] + normals[k][1] * delta)))
                dest.append(src[j])
                dest.append((_cround(src[j][0] + normals[j][0] * delta), _cround(src[j][1] + normals[j][1] * delta)))
            else:  # DoRound
                a = math.atan2(sin_a, normals[k][0] * normals[j][0] + normals[k][1] * normals[j][1])
                st = max(int(_cround(steps_per_rad * abs(a))), 1)
                X, Y = normals[k]
                for _ in range(st):
                    dest.append((_cround(src[j][0] + X * delta), _cround(src[j][1] + Y * delta)))
                    X2 = X
                    X = X * m_cos - m_sin * Y
                    Y = X2 * m_sin + Y * m_cos
                dest.append((_cround(src[j][0] + normals[j][0] * delta), _cround(src[j][1] + normals[j][1] * delta)))
        k = j
    return dest


class PyclipperStandIn:
    """Just enough of the pyclipper API for DBPostProcess.unclip; injected where the real wheel is missing."""
    JT_ROUND = 1
    ET_CLOSEDPOLYGON = 0

    class PyclipperOffset:
        def __init__(self, miter_limit=2.0, arc_tolerance=0.25):
            self.arc_tolerance = arc_tolerance
            self.path = None

        def AddPath(self, path, join_type, end_type):
            self.path = [(int(p[0]), int(p[1])) for p in path]  # pyclipper: <cInt> cast = truncation

        def Execute(self, delta):
            out = clipper_offset_round(self.path, float(delta), self.arc_tolerance)
            return [[list(p) for p in out]] if out else []


class ShapelyPolygonStandIn:
    """shapely.geometry.Polygon(box).area / .length for a simple polygon (float64)."""

    def __init__(self, pts):
        self.p = np.asarray(pts, np.float64).reshape(-1, 2)

    @property
    def area(self):
        x, y = self.p[:, 0], self.p[:, 1]
        return abs(float(np.dot(x, np.roll(y, -1)) - np.dot(y, np.roll(x, -1)))) * 0.5

    @property
    def length(self):
        d = self.p - np.roll(self.p, -1, axis=0)
        return float(np.sqrt((d * d).sum(1)).sum())


# ------------------------------------------------------------------------------------------- DBPostProcess restated
def get_mini_boxes(contour):
    bounding_box = cv2.minAreaRect(contour)
    points = sorted(list(cv2.boxPoints(bounding_box)), key=lambda x: x[0])
    if points[1][1] > points[0][1]:
        index_1, index_4 = 0, 1
    else:
        index_1, index_4 = 1, 0
    if points[3][1] > points[2][1]:
        index_2, index_3 = 2, 3
    else:
        index_2, index_3 = 3, 2
    return [points[index_1], points[index_2], points[index_3], points[index_4]], min(bounding_box[1])


def box_score_fast(bitmap, _box):
    h, w = bitmap.shape[:2]
    box = _box.copy()
    xmin = np.clip(np.floor(box[:, 0].min()).astype(int), 0, w - 1)
    xmax = np.clip(np.ceil(box[:, 0].max()).astype(int), 0, w - 1)
    ymin = np.clip(np.floor(box[:, 1].min()).astype(int), 0, h - 1)
    ymax = np.clip(np.ceil(box[:, 1].max()).astype(int), 0, h - 1)
    mask = np.zeros((ymax - ymin + 1, xmax - xmin + 1), dtype=np.uint8)
    box[:, 0] = box[:, 0] - xmin
    box[:, 1] = box[:, 1] - ymin
    cv2.fillPoly(mask, box.reshape(1, -1, 2).astype(np.int32), 1)
    return cv2.mean(bitmap[ymin:ymax + 1, xmin:xmax + 1], mask)[0]


def unclip(box, unclip_ratio):
    poly = ShapelyPolygonStandIn(box)
    distance = poly.area * unclip_ratio / poly.length
    offset = PyclipperStandIn.PyclipperOffset()
    offset.AddPath(box, PyclipperStandIn.JT_ROUND, PyclipperStandIn.ET_CLOSEDPOLYGON)
    return np.array(offset.Execute(distance))


def boxes_from_bitmap(pred, bitmap, dest_width, dest_height, box_thresh=0.6, unclip_ratio=1.5, max_candidates=1000,
                      min_size=3, return_scores=False):
    height, width = bitmap.shape
    contours, _ = cv2.findContours((bitmap * 255).astype(np.uint8), cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    boxes, scores = [], []
    for index in range(min(len(contours), max_candidates)):
        points, sside = get_mini_boxes(contours[index])
        if sside < min_size:
            continue
        points = np.array(points)
        score = box_score_fast(pred, points.reshape(-1, 2))
        if box_thresh > score:
            continue
        box = unclip(points, unclip_ratio).reshape(-1, 1, 2)
        box, sside = get_mini_boxes(box)
        if sside < min_size + 2:
            continue
        box = np.array(box)
        box[:, 0] = np.clip(np.round(box[:, 0] / width * dest_width), 0, dest_width)
        box[:, 1] = np.clip(np.round(box[:, 1] / height * dest_height), 0, dest_height)
        boxes.append(box.astype(np.int16))
        scores.append(score)
    return np.array(boxes, dtype=np.int16), scores


def order_points_clockwise(pts):
    x_sorted = pts[np.argsort(pts[:, 0]), :]
    left, right = x_sorted[:2, :], x_sorted[2:, :]
    left = left[np.argsort(left[:, 1]), :]
    (tl, bl) = left
    right = right[np.argsort(right[:, 1]), :]
    (tr, br) = right
    return np.array([tl, tr, br, bl], dtype="float32")


def filter_tag_det_res(dt_boxes, image_shape):
    img_height, img_width = image_shape[0:2]
    out = []
    for box in dt_boxes:
        box = order_points_clockwise(box)
        for pno in range(box.shape[0]):
            box[pno, 0] = int(min(max(box[pno, 0], 0), img_width - 1))
            box[pno, 1] = int(min(max(box[pno, 1], 0), img_height - 1))
        rect_width = int(np.linalg.norm(box[0] - box[1]))
        rect_height = int(np.linalg.norm(box[0] - box[3]))
        if rect_width <= 3 or rect_height <= 3:
            continue
        out.append(box)
    return np.array(out)


def db_postprocess(pred: np.ndarray, shape_list: np.ndarray, org_shape, thresh=0.2, box_thresh=0.6, unclip_ratio=1.5,
                   max_candidates=1000) -> np.ndarray:
    """pred fp32 [H,W] probability map of ONE page; shape_list = np.array([src_h, src_w, ratio_h, ratio_w]) (float64);
    org_shape = (h, w[, c]) -> det_polygons float32 [n, 8]  (PPOcrDetectionPostProcessor.__call__)."""
    pred = np.asarray(pred, np.float32)
    seg = pred > thresh
    src_h, src_w = shape_list[0], shape_list[1]
    boxes, _ = boxes_from_bitmap(pred, seg, src_w, src_h, box_thresh, unclip_ratio, max_candidates)
    return filter_tag_det_res(boxes, org_shape).reshape(-1, 8)


def dbnet_boxes_from_bitmap(pred, bitmap, dest_width, dest_height, box_thresh=0.3, unclip_ratio=1.5, max_candidates=1000):
    """boxes_from_bitmap of the in-tree DBNet back-end (db_net/ocr_detection_utils.py:168-205): as the db_pp one up to the second
    mini box, then the box is TRUNCATED to int32, scaled with np.round(box / width * dest) in float64, clipped to [0, dest]
    and returned as a flat list of python ints; no clockwise re-ordering / size filter follows."""
    height, width = bitmap.shape
    contours, _ = cv2.findContours((bitmap * 255).astype(np.uint8), cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)
    boxes, scores = [], []
    for contour in contours[:max_candidates]:
        points, sside = get_mini_boxes(contour)
        if sside < 3:
            continue
        points = np.array(points)
        score = box_score_fast(pred, points.reshape(-1, 2))
        if box_thresh > score:
            continue
        box = unclip(points, unclip_ratio).reshape(-1, 1, 2)
        box, sside = get_mini_boxes(box)
        if sside < 3 + 2:
            continue
        box = np.array(box).astype(np.int32)
        box[:, 0] = np.clip(np.round(box[:, 0] / width * dest_width), 0, dest_width)
        box[:, 1] = np.clip(np.round(box[:, 1] / height * dest_height), 0, dest_height)
        boxes.append(box.reshape(-1).tolist())
        scores.append(score)
    return boxes, scores


def dbnet_postprocess(pred: np.ndarray, org_shape, thresh=0.2) -> np.ndarray:
    """OCRDetectionPostProcessor.__call__ (db_net/processor_ocr_dbnet.py:113-127): pred fp32 [H,W] of one page, org_shape =
    (height, width) -> det_polygons int64 [n, 8]."""
    pred = np.asarray(pred, np.float32)
    height, width = int(org_shape[0]), int(org_shape[1])
    boxes, _ = dbnet_boxes_from_bitmap(pred, pred > thresh, width, height)
    return np.array(boxes, dtype=np.int64).reshape(-1, 8)
