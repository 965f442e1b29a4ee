"""Prototype: exact restatement of cv2.fillPoly (8-connected, shift 0) for a convex quad with integer vertices,
checked against cv2 (used to design the mask test of the db_boxes kernel)."""
import numpy as np, cv2

XY_SHIFT = 16
XY_ONE = 1 << 16


def line_pixels(p1, p2):
    """cv::LineIterator(connectivity 8, leftToRight=true) pixel list."""
    x1, y1 = p1
    x2, y2 = p2
    dx, dy = x2 - x1, y2 - y1
    if dx < 0:  # left to right: swap endpoints
        x1, y1, x2, y2 = x2, y2, x1, y1
        dx, dy = -dx, -dy
    ystep = 1
    if dy < 0:
        dy = -dy
        ystep = -1
    pts = []
    if dy > dx:  # steep: step along y
        err = dy - 2 * dx
        plus, minus = 2 * dy, -2 * dx
        x, y = x1, y1
        for _ in range(dy + 1):
            pts.append((x, y))
            if err < 0:
                err += minus + plus
                x += 1
                y += ystep
            else:
                err += minus
                y += ystep
    else:
        err = dx - 2 * dy
        plus, minus = 2 * dx, -2 * dy
        x, y = x1, y1
        for _ in range(dx + 1):
            pts.append((x, y))
            if err < 0:
                err += minus + plus
                x += 1
                y += ystep
            else:
                err += minus
                x += 1
    return pts


def cdiv(a, b):
    """C++ integer division (truncation toward zero)."""
    q = abs(a) // abs(b)
    return q if (a >= 0) == (b >= 0) else -q


def fill_quad(V, H, W, half=True):
    mask = np.zeros((H, W), np.uint8)
    edges = []
    n = len(V)
    for i in range(n):
        p0, p1 = V[i - 1], V[i]
        for (x, y) in line_pixels(tuple(p0), tuple(p1)):
            if 0 <= x < W and 0 <= y < H:
                mask[y, x] = 1
        if p0[1] == p1[1]:
            continue
        x0 = (int(p0[0]) << XY_SHIFT) + (XY_ONE >> 1 if half else 0)
        x1 = (int(p1[0]) << XY_SHIFT) + (XY_ONE >> 1 if half else 0)
        dxf = cdiv(x1 - x0, int(p1[1]) - int(p0[1]))
        if p0[1] < p1[1]:
            edges.append([int(p0[1]), int(p1[1]), x0, dxf])
        else:
            edges.append([int(p1[1]), int(p0[1]), x1, dxf])
    if len(edges) < 2:
        return mask
    ymin = min(e[0] for e in edges)
    ymax = min(max(e[1] for e in edges), H)
    for y in range(ymin, ymax):
        xs = sorted(e[2] + (y - e[0]) * e[3] for e in edges if e[0] <= y < e[1])
        for k in range(0, len(xs) - 1, 2):
            a, b = xs[k] >> XY_SHIFT, xs[k + 1] >> XY_SHIFT
            if a < W and b >= 0:
                mask[y, max(a, 0):min(b, W - 1) + 1] = 1
    return mask


if __name__ == "__main__":
    rng = np.random.default_rng(1)
    for half in (True, False):
        bad = 0
        for trial in range(4000):
            cx, cy = rng.uniform(20, 60, 2)
            w, h = rng.uniform(3, 60), rng.uniform(3, 25)
            ang = rng.uniform(-90, 90) if trial % 3 else 0.0
            pts = cv2.boxPoints(((cx, cy), (w, h), ang))
            pts = sorted(list(pts), key=lambda p: p[0])
            i1, i4 = (0, 1) if pts[1][1] > pts[0][1] else (1, 0)
            i2, i3 = (2, 3) if pts[3][1] > pts[2][1] else (3, 2)
            box = np.array([pts[i1], pts[i2], pts[i3], pts[i4]], np.float32)
            box -= np.floor(box.min(0))
            V = box.astype(np.int32)
            H, W = int(V[:, 1].max()) + 2, int(V[:, 0].max()) + 2
            ref = np.zeros((H, W), np.uint8)
            cv2.fillPoly(ref, V.reshape(1, -1, 2), 1)
            got = fill_quad(V, H, W, half)
            if (ref != got).any():
                bad += 1
                if bad < 3:
                    print(V.tolist()); print(ref.astype(int) - got.astype(int))
        print("half", half, "mismatching quads:", bad)


def fill_quad_var(V, H, W, common, ldelta, rdelta, lines=True):
    mask = np.zeros((H, W), np.uint8)
    edges = []
    for i in range(len(V)):
        p0, p1 = V[i - 1], V[i]
        if lines:
            for (x, y) in line_pixels(tuple(p0), tuple(p1)):
                if 0 <= x < W and 0 <= y < H:
                    mask[y, x] = 1
        if p0[1] == p1[1]:
            continue
        x0 = (int(p0[0]) << XY_SHIFT) + common
        x1 = (int(p1[0]) << XY_SHIFT) + common
        dxf = cdiv(x1 - x0, int(p1[1]) - int(p0[1]))
        if p0[1] < p1[1]:
            edges.append([int(p0[1]), int(p1[1]), x0, dxf])
        else:
            edges.append([int(p1[1]), int(p0[1]), x1, dxf])
    if len(edges) < 2:
        return mask
    ymin = min(e[0] for e in edges)
    ymax = min(max(e[1] for e in edges), H)
    for y in range(ymin, ymax):
        xs = sorted(e[2] + (y - e[0]) * e[3] for e in edges if e[0] <= y < e[1])
        for k in range(0, len(xs) - 1, 2):
            a, b = (xs[k] + ldelta) >> XY_SHIFT, (xs[k + 1] + rdelta) >> XY_SHIFT
            if a <= b and a < W and b >= 0:
                mask[y, max(a, 0):min(b, W - 1) + 1] = 1
    return mask
