#!/usr/bin/env python
"""How close is the restatement of cv::minAreaRect (oracle/cv_geom_ref.py, the numpy twin of csrc/db_post.cu) to the cv2 of
this image, bit for bit?  Random filled quads / discs -> findContours -> cv2.minAreaRect vs min_area_rect.

    python tools/min_area_rect_probe.py [n]

Findings with OpenCV 4.13 (recorded in DESIGN.md): the calipers walk the COUNTER-clockwise hull, choose the next edge by
cross-product signs (firstVecIsRight) instead of cosines, and the angle is normalised into [-90, 0) in double.  With the three
restated every random contour agrees with cv2 in every bit of (centre, size, angle); round 1's restatement (clockwise hull,
cosine comparison, float angle shift) agreed on ~10 % of rotated rectangles."""
import os
import sys

import cv2
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import cv_geom_ref as G  # noqa: E402


def contours(n, seed=11):
    rng = np.random.default_rng(seed)
    for t in range(n):
        img = np.zeros((220, 320), np.uint8)
        cx, cy = rng.uniform(80, 240), rng.uniform(80, 140)
        w, h, a = rng.uniform(2, 120), rng.uniform(2, 120), rng.uniform(-180, 180)
        if t % 5 == 0:
            a = float(rng.choice([0, 90, 45, -45, 30]))
        cv2.fillPoly(img, [cv2.boxPoints(((cx, cy), (w, h), a)).astype(np.int32)], 1)
        if t % 7 == 0:
            cv2.circle(img, (int(cx), int(cy)), int(rng.uniform(3, 40)), 1, -1)
        for c in cv2.findContours(img, cv2.RETR_LIST, cv2.CHAIN_APPROX_SIMPLE)[0]:
            yield c


def compare(n=800, seed=11):
    """(contours, bit-identical, identical after the int() truncation the DB post-process applies to boxPoints, worst |diff|)."""
    tot = same = same_int = 0
    worst = 0.0
    for c in contours(n, seed):
        r = cv2.minAreaRect(c)
        m = G.min_area_rect([(int(p[0]), int(p[1])) for p in c.reshape(-1, 2)])
        rv = np.array([r[0][0], r[0][1], r[1][0], r[1][1], r[2]], np.float32)
        mv = np.array([m[0][0], m[0][1], m[1][0], m[1][1], m[2]], np.float32)
        tot += 1
        same += bool((rv.view(np.uint32) == mv.view(np.uint32)).all())
        worst = max(worst, float(np.abs(rv - mv).max()))
        same_int += bool((cv2.boxPoints(r).astype(np.int32) == G.box_points(m).astype(np.int32)).all())
    return tot, same, same_int, worst


if __name__ == "__main__":
    tot, same, same_int, worst = compare(int(sys.argv[1]) if len(sys.argv) > 1 else 800)
    print(f"{tot} contours: {same} bit-identical ({same / tot:.3f}), {same_int} identical after int truncation of boxPoints ({same_int / tot:.4f}), worst |diff| {worst:.3g}")
