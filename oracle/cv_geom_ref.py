"""Pure-Python restatement of the OpenCV routines on the DB post-process path (TEST ORACLE / kernel design aid):
border following (findContours RETR_LIST + CHAIN_APPROX_SIMPLE), convexHull (Sklansky), minAreaRect (rotating
calipers) and boxPoints, written to be BIT-EXACT with cv2 (float32 arithmetic where OpenCV uses float, double where
it uses double), because the reference truncates these float corners to integers before the Clipper offset
(db_pp/processor_ocr_db_pp.py:221-228 -> pyclipper <cInt> cast): a last-bit difference moves a box by one pixel.
Checked against cv2 itself in tests/test_oracle_cpu.py; the CUDA kernel (csrc/db_post.cu) mirrors this file.
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32
# 8-neighbourhood codes of OpenCV (icvCodeDeltas): 0=E, 1=NE, 2=N, 3=NW, 4=W, 5=SW, 6=S, 7=SE (y grows downwards)
CODE_DX = (1, 1, 0, -1, -1, -1, 0, 1)
CODE_DY = (0, -1, -1, -1, 0, 1, 1, 1)


def trace_border(img: np.ndarray, x0: int, y0: int, is_hole: bool):
    """icvFetchContour with CHAIN_APPROX_SIMPLE on a binary image (nonzero = foreground; a zero frame is implied).
    (x0, y0) is the start pixel found by the raster scan: the top-left pixel of an outer border, or the pixel left of
    the first background pixel of a hole.  Returns the list of vertices (x, y)."""
    h, w = img.shape

    def px(x, y):
        return 0 <= x < w and 0 <= y < h and img[y, x] != 0

    s_end = s = 0 if is_hole else 4
    while True:
        s = (s - 1) & 7
        if px(x0 + CODE_DX[s], y0 + CODE_DY[s]) or s == s_end:
            break
    if s == s_end and not px(x0 + CODE_DX[s], y0 + CODE_DY[s]):
        return [(x0, y0)]
    if s == s_end:  # loop ended on s_end with a foreground neighbour: OpenCV tests `*i1 == 0 && s != s_end`
        pass
    i1 = (x0 + CODE_DX[s], y0 + CODE_DY[s])
    out = []
    x3, y3 = x0, y0
    prev_s = s ^ 4
    ptx, pty = x0, y0
    while True:
        s_end = s
        while True:  # search the next border pixel counter-clockwise from s+1
            s += 1
            x4, y4 = x3 + CODE_DX[s & 7], y3 + CODE_DY[s & 7]
            if px(x4, y4):
                break
        s &= 7
        if s != prev_s:
            out.append((ptx, pty))
            prev_s = s
        ptx += CODE_DX[s]
        pty += CODE_DY[s]
        if (x4, y4) == (x0, y0) and (x3, y3) == i1:
            break
        x3, y3 = x4, y4
        s = (s + 4) & 7
    return out


def _sklansky(pts, start, end, nsign, sign2):
    """convhull.cpp Sklansky_ on pointer-sorted integer points; returns the stack (indices into pts)."""
    incr = 1 if end > start else -1
    pprev, pcur = start, start + incr
    pnext = pcur + incr
    if start == end or (pts[start][0] == pts[end][0] and pts[start][1] == pts[end][1]):
        return [start]
    stack = [pprev, pcur, pnext]
    end += incr
    sign = lambda v: (v > 0) - (v < 0)
    while pnext != end:
        cury, nexty = pts[pcur][1], pts[pnext][1]
        by = nexty - cury
        if sign(by) != nsign:
            ax = pts[pcur][0] - pts[pprev][0]
            bx = pts[pnext][0] - pts[pcur][0]
            ay = cury - pts[pprev][1]
            convexity = ay * bx - ax * by
            if sign(convexity) == sign2 and (ax != 0 or ay != 0):
                pprev, pcur = pcur, pnext
                pnext += incr
                stack.append(pnext)
            else:
                if pprev == start:
                    pcur = pnext
                    stack[1] = pcur
                    pnext += incr
                    stack[2] = pnext
                else:
                    stack[-2] = pnext
                    pcur = pprev
                    pprev = stack[-4]
                    stack.pop()
        else:
            pnext += incr
            stack[-1] = pnext
    stack.pop()
    return stack


def convex_hull(points, clockwise=False):
    """cv::convexHull(points, clockwise, returnPoints=true) for integer points; list of (x, y)."""
    total = len(points)
    if total == 0:
        return []
    order = sorted(range(total), key=lambda i: (points[i][0], points[i][1], i))
    pts = [points[i] for i in order]
    miny_ind = maxy_ind = 0
    for i in range(1, total):
        y = pts[i][1]
        if pts[miny_ind][1] > y:
            miny_ind = i
        if pts[maxy_ind][1] < y:
            maxy_ind = i
    hullbuf = []
    if pts[0] == pts[total - 1]:
        hullbuf.append(0)
    else:
        tl = _sklansky(pts, 0, maxy_ind, -1, 1)
        tr = _sklansky(pts, total - 1, maxy_ind, -1, -1)
        if not clockwise:
            tl, tr = tr, tl
        for i in range(len(tl) - 1):
            hullbuf.append(tl[i])
        for i in range(len(tr) - 1, 0, -1):
            hullbuf.append(tr[i])
        stop_idx = tr[1] if len(tr) > 2 else (tl[len(tl) - 2] if len(tl) > 2 else -1)
        bl = _sklansky(pts, 0, miny_ind, 1, -1)
        br = _sklansky(pts, total - 1, miny_ind, 1, 1)
        if clockwise:
            bl, br = br, bl
        if stop_idx >= 0:
            check_idx = bl[1] if len(bl) > 2 else (br[2 - len(bl)] if len(bl) + len(br) > 2 else -1)
            if check_idx == stop_idx or (check_idx >= 0 and pts[check_idx] == pts[stop_idx]):
                bl = bl[:min(len(bl), 2)]
                br = br[:min(len(br), 2)]
        for i in range(len(bl) - 1):
            hullbuf.append(bl[i])
        for i in range(len(br) - 1, 0, -1):
            hullbuf.append(br[i])
    # map back to original indices, then the cyclic shift that makes original indices monotone
    hull_idx = [order[i] for i in hullbuf]
    nout = len(hull_idx)
    if nout >= 3:
        min_idx = max_idx = lt = 0
        broke = False
        for i in range(1, nout):
            idx = hull_idx[i]
            lt += hull_idx[i - 1] < idx
            if lt > 1 and lt <= i - 2:
                broke = True
                break
            if idx < hull_idx[min_idx]:
                min_idx = i
            if idx > hull_idx[max_idx]:
                max_idx = i
        mmdist = abs(max_idx - min_idx)
        if (mmdist == 1 or mmdist == nout - 1) and (lt <= 1 or lt >= nout - 2):
            ascending = (max_idx + 1) % nout == min_idx
            i0 = min_idx if ascending else max_idx
            j = i0
            if i0 > 0:
                tmp = []
                ok = True
                for i in range(nout):
                    curr = hull_idx[j]
                    tmp.append(curr)
                    next_j = j + 1 if j + 1 < nout else 0
                    nxt = hull_idx[next_j]
                    if i < nout - 1 and (ascending != (curr < nxt)):
                        ok = False
                        break
                    j = next_j
                if ok:
                    hull_idx = tmp
    return [points[i] for i in hull_idx]


def rotating_calipers(hp):
    """rotcalipers.cpp rotatingCalipers(CALIPERS_MINAREARECT) on float32 hull points; returns out[6] (float32)."""
    n = len(hp)
    px = [F(p[0]) for p in hp]
    py = [F(p[1]) for p in hp]
    vx, vy, inv = [F(0)] * n, [F(0)] * n, [F(0)] * n
    left = bottom = right = top = 0
    left_x = right_x = px[0]
    top_y = bottom_y = py[0]
    p0x, p0y = px[0], py[0]
    for i in range(n):
        if p0x < left_x:
            left_x, left = p0x, i
        if p0x > right_x:
            right_x, right = p0x, i
        if p0y > top_y:
            top_y, top = p0y, i
        if p0y < bottom_y:
            bottom_y, bottom = p0y, i
        j = i + 1 if i + 1 < n else 0
        dx = float(px[j]) - float(p0x)
        dy = float(py[j]) - float(p0y)
        vx[i], vy[i] = F(dx), F(dy)
        inv[i] = F(1.0 / math.sqrt(dx * dx + dy * dy))
        p0x, p0y = px[j], py[j]
    orientation = F(0)
    ax, ay = float(vx[n - 1]), float(vy[n - 1])
    for i in range(n):
        bx, by = float(vx[i]), float(vy[i])
        convexity = ax * by - ay * bx
        if convexity != 0:
            orientation = F(1) if convexity > 0 else F(-1)
            break
        ax, ay = bx, by
    base_a, base_b = orientation, F(0)
    seq = [bottom, right, top, left]
    minarea = F(np.finfo(np.float32).max)
    buf = None
    for k in range(n):
        # OpenCV >= 4.5.2: the caliper side with the smallest angle to its polygon edge is found by cross-product signs of the
        # edge vectors rotated into a common frame (firstVecIsRight), not by comparing cosines
        rot = [(vx[seq[0]], vy[seq[0]]), (vy[seq[1]], F(-vx[seq[1]])), (F(-vx[seq[2]]), F(-vy[seq[2]])), (F(-vy[seq[3]]), vx[seq[3]])]
        main = 0
        for i in range(1, 4):
            tx, ty = rot[i][1], F(-rot[i][0])  # rotate90CW(rot[i])
            if F(F(tx * rot[main][0]) + F(ty * rot[main][1])) < 0:
                main = i
        pi = seq[main]
        lead_x = F(vx[pi] * inv[pi])
        lead_y = F(vy[pi] * inv[pi])
        if main == 0:
            base_a, base_b = lead_x, lead_y
        elif main == 1:
            base_a, base_b = lead_y, F(-lead_x)
        elif main == 2:
            base_a, base_b = F(-lead_x), F(-lead_y)
        else:
            base_a, base_b = F(-lead_y), lead_x
        seq[main] += 1
        if seq[main] == n:
            seq[main] = 0
        dx = F(px[seq[1]] - px[seq[3]])
        dy = F(py[seq[1]] - py[seq[3]])
        width = F(F(dx * base_a) + F(dy * base_b))
        dx = F(px[seq[2]] - px[seq[0]])
        dy = F(py[seq[2]] - py[seq[0]])
        height = F(F(-dx * base_b) + F(dy * base_a))
        area = F(width * height)
        if area <= minarea:
            minarea = area
            buf = (seq[3], base_a, width, base_b, height, seq[0])
    l_i, A1, w_, B1, h_, b_i = buf
    A2, B2 = F(-B1), A1
    C1 = F(F(A1 * px[l_i]) + F(py[l_i] * B1))
    C2 = F(F(A2 * px[b_i]) + F(py[b_i] * B2))
    idet = F(F(1) / F(F(A1 * B2) - F(A2 * B1)))
    ox = F(F(F(C1 * B2) - F(C2 * B1)) * idet)
    oy = F(F(F(A1 * C2) - F(A2 * C1)) * idet)
    return [ox, oy, F(A1 * w_), F(B1 * w_), F(A2 * h_), F(B2 * h_)]


def min_area_rect(points):
    """cv::minAreaRect of OpenCV 4.13 for integer points -> ((cx, cy), (w, h), angle) as float32 (angle in degrees).
    Two facts measured against cv2 4.13 (tools/min_area_rect_probe.py; pinned by tests/test_oracle_cpu.py): the calipers walk
    the COUNTER-clockwise hull, and the angle is brought into [-90, 0) in double by quarter turns that swap width and height,
    then rounded to float once."""
    hull = convex_hull(points, clockwise=False)
    n = len(hull)
    cx = cy = w = h = F(0)
    ang = 0.0
    if n > 2:
        o = rotating_calipers(hull)
        cx = F(o[0] + F(F(o[2] + o[4]) * F(0.5)))
        cy = F(o[1] + F(F(o[3] + o[5]) * F(0.5)))
        w = F(math.sqrt(float(o[2]) * float(o[2]) + float(o[3]) * float(o[3])))
        h = F(math.sqrt(float(o[4]) * float(o[4]) + float(o[5]) * float(o[5])))
        ang = math.atan2(float(o[3]), float(o[2]))
    elif n == 2:
        cx = F(F(F(hull[0][0]) + F(hull[1][0])) * F(0.5))
        cy = F(F(F(hull[0][1]) + F(hull[1][1])) * F(0.5))
        dx = float(F(hull[1][0]) - F(hull[0][0]))
        dy = float(F(hull[1][1]) - F(hull[0][1]))
        w = F(math.sqrt(dx * dx + dy * dy))
        ang = math.atan2(dy, dx)
    elif n == 1:
        cx, cy = F(hull[0][0]), F(hull[0][1])
    ang = ang * 180 / math.pi
    while ang >= 0.0:
        ang -= 90.0
        w, h = h, w
    while ang < -90.0:
        ang += 90.0
        w, h = h, w
    return (cx, cy), (w, h), F(ang)


def box_points(rect):
    """cv::RotatedRect::points."""
    (cx, cy), (w, h), ang = rect
    a_ = float(ang) * math.pi / 180.0
    b = F(F(math.cos(a_)) * F(0.5))
    a = F(F(math.sin(a_)) * F(0.5))
    p0x = F(F(cx - F(a * h)) - F(b * w))
    p0y = F(F(cy + F(b * h)) - F(a * w))
    p1x = F(F(cx + F(a * h)) - F(b * w))
    p1y = F(F(cy - F(b * h)) - F(a * w))
    p2x, p2y = F(F(F(2) * cx) - p0x), F(F(F(2) * cy) - p0y)
    p3x, p3y = F(F(F(2) * cx) - p1x), F(F(F(2) * cy) - p1y)
    return np.array([[p0x, p0y], [p1x, p1y], [p2x, p2y], [p3x, p3y]], np.float32)
