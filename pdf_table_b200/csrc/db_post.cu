// db_threshold_seed: DB probability map -> text boxes, entirely on the GPU.
//
// Follows the reference DBPostProcess / PPOcrDetectionPostProcessor (db_pp/processor_ocr_db_pp.py:174-311, 330-386):
//   pred > thresh -> cv2.findContours(RETR_LIST, CHAIN_APPROX_SIMPLE) -> first max_candidates contours ->
//   get_mini_boxes (cv2.minAreaRect + boxPoints) -> sside < 3 -> box_score_fast (cv2.fillPoly mask, cv2.mean) ->
//   < box_thresh -> unclip (shapely area/length, pyclipper round offset) -> get_mini_boxes -> sside < 5 ->
//   scale / round / clip / int16 -> order_points_clockwise -> clip -> drop w/h <= 3.
//
// GPU formulation
//   1. k_db_label_init/merge/flatten: one union-find labelling of BOTH classes at once -- foreground with
//      8-connectivity, background with 4-connectivity (that is the topology of Suzuki border following); background
//      components that touch the frame are the outside, every other background component is a hole.
//   2. k_db_enumerate/rank: a contour exists per foreground component (outer border, start = its raster-first
//      pixel) and per hole (hole border, start = the pixel left of the hole's raster-first pixel).  cv2 returns
//      contours in DESCENDING raster order of those start pixels, so rank = number of contours with a larger start
//      key; contours with rank < max_candidates get slot = rank (that is the reference's "first N contours").
//   3. k_db_contour_boxes: one warp per slot.  Lane 0 re-traces the border exactly as icvFetchContour does (the
//      contour's vertex ORDER feeds cv::convexHull's index-monotone cyclic shift and hence the float32 rounding of
//      cv::minAreaRect, which the reference then TRUNCATES to integers for Clipper), the warp sorts the vertices,
//      lane 0 runs Sklansky + rotating calipers + boxPoints in the same float32/float64 operations as OpenCV, the
//      warp evaluates the fillPoly mask mean (closed-form Bresenham edges + 16.16 fixed-point scan-line spans), and
//      lane 0 finishes with the Clipper round offset, the second minAreaRect and the integer box arithmetic.
//   4. k_db_compact: boxes are emitted in slot (= reference) order, left-packed per page.
// The mirror of the geometry in plain Python is oracle/cv_geom_ref.py (checked against cv2) and oracle/db_post_ref.py.
#include <float.h>
#include <math.h>

#include "engine.h"

namespace dv {

namespace {

constexpr int kMaxSlots = 1000;  // hard cap of boxes per page (the reference's max_candidates default)
constexpr int kMaxV = 2048;      // contour vertices (after CHAIN_APPROX_SIMPLE) handled per contour
constexpr int kMaxOff = 1024;    // vertices of a Clipper round-offset polygon

struct DbWs : Model {
    int N = 0, H = 0, W = 0;
    int* label = nullptr;
    uint8_t* fg = nullptr;
    uint8_t* outer = nullptr;
    int* rowcnt = nullptr;
    int* rowsuf = nullptr;
    int* ncont = nullptr;
    int* slot_root = nullptr;  // [N][kMaxSlots] root pixel index (page-local), -1 = empty
    float* slot_box = nullptr;  // [N][kMaxSlots][8]
    int* slot_valid = nullptr;  // [N][kMaxSlots]
    double* src_hw = nullptr;   // [N][2]
    int* overflow = nullptr;    // [1] contours skipped because they exceed kMaxV
    std::vector<void*> mem;
    ~DbWs() override {
        for (void* p : mem) cudaFree(p);
    }
};

// ------------------------------------------------------------------------------------------------ labelling
__device__ __forceinline__ int find_root(const int* L, int a) {
    int l = reinterpret_cast<const volatile int*>(L)[a];
    while (l != a) {
        a = l;
        l = reinterpret_cast<const volatile int*>(L)[a];
    }
    return a;
}
__device__ __forceinline__ void unite(int* L, int a, int b) {
    bool done;
    do {
        a = find_root(L, a);
        b = find_root(L, b);
        if (a < b) {
            const int old = atomicMin(&L[b], a);
            done = (old == b);
            b = old;
        } else if (b < a) {
            const int old = atomicMin(&L[a], b);
            done = (old == a);
            a = old;
        } else {
            done = true;
        }
    } while (!done);
}

// grid: (ceil(W/256), H, N), block 256: a warp covers 32 consecutive pixels of one row
__global__ void __launch_bounds__(256)
k_db_label_init(const float* __restrict__ prob, int H, int W, float thresh, int* __restrict__ label, uint8_t* __restrict__ fg) {
    const int x = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y, n = blockIdx.z;
    const int lane = threadIdx.x & 31;
    const long long base = (static_cast<long long>(n) * H + y) * W;
    const bool in = x < W;
    int c = 2;
    if (in) c = prob[base + x] > thresh ? 1 : 0;  // float32 compare, as numpy does with a python-float threshold
    const int prev = __shfl_up_sync(0xffffffffu, c, 1);
    const bool boundary = (lane == 0) || (prev != c);
    const unsigned m = __ballot_sync(0xffffffffu, boundary);
    const int start = 31 - __clz(m & (0xffffffffu >> (31 - lane)));
    if (in) {
        fg[base + x] = static_cast<uint8_t>(c);
        label[base + x] = static_cast<int>(base) + x - lane + start;
    }
}

__global__ void __launch_bounds__(256)
k_db_label_merge(int H, int W, int* __restrict__ label, const uint8_t* __restrict__ fg) {
    const int x = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y, n = blockIdx.z;
    if (x >= W) return;
    const int p = static_cast<int>((static_cast<long long>(n) * H + y) * W) + x;
    const int c = fg[p];
    const bool left = x > 0 && fg[p - 1] == c;
    if (left && (threadIdx.x & 31) == 0) unite(label, p, p - 1);  // run continues across a warp boundary
    if (y == 0) return;
    const bool up = fg[p - W] == c;
    if (c) {  // foreground: 8-connectivity; only the links a run-based scan would not already imply
        if (up) {
            if (!left) unite(label, p, p - W);
        } else if (x > 0 && fg[p - W - 1] && !left) {
            unite(label, p, p - W - 1);
        }
        if (x < W - 1 && fg[p - W + 1] && !up) unite(label, p, p - W + 1);
    } else {  // background: 4-connectivity
        if (up) {
            const bool upl = x > 0 && fg[p - W - 1] == 0;
            if (!(left && upl)) unite(label, p, p - W);
        }
    }
}

__global__ void __launch_bounds__(256)
k_db_label_flatten(int H, int W, int* __restrict__ label, const uint8_t* __restrict__ fg, uint8_t* __restrict__ outer) {
    const int x = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y, n = blockIdx.z;
    if (x >= W) return;
    const int p = static_cast<int>((static_cast<long long>(n) * H + y) * W) + x;
    const int r = find_root(label, p);
    label[p] = r;
    if (!fg[p] && (x == 0 || y == 0 || x == W - 1 || y == H - 1)) outer[r] = 1;
}

// contour roots: foreground roots, and background roots that are not the outside.  key row = row of the start pixel.
__global__ void __launch_bounds__(256)
k_db_enumerate(int H, int W, const int* __restrict__ label, const uint8_t* __restrict__ fg, const uint8_t* __restrict__ outer,
               int* __restrict__ rowcnt, int* __restrict__ ncont) {
    const int x = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y, n = blockIdx.z;
    if (x >= W) return;
    const int p = static_cast<int>((static_cast<long long>(n) * H + y) * W) + x;
    if (label[p] != p) return;
    if (!fg[p] && outer[p]) return;
    atomicAdd(&rowcnt[n * (H + 1) + y], 1);
    atomicAdd(&ncont[n], 1);
}

__global__ void k_db_rowsuffix(int H, const int* __restrict__ rowcnt, int* __restrict__ rowsuf) {
    const int n = blockIdx.x;
    if (threadIdx.x != 0) return;
    int acc = 0;
    for (int y = H; y >= 0; --y) {  // rowsuf[y] = number of contours starting in rows > y
        rowsuf[n * (H + 1) + y] = acc;
        if (y <= H - 1) acc += rowcnt[n * (H + 1) + y];
    }
}

__device__ __forceinline__ bool is_contour_root(const int* label, const uint8_t* fg, const uint8_t* outer, int p) {
    return label[p] == p && (fg[p] || !outer[p]);
}

__global__ void __launch_bounds__(256)
k_db_rank(int H, int W, const int* __restrict__ label, const uint8_t* __restrict__ fg, const uint8_t* __restrict__ outer,
          const int* __restrict__ rowsuf, int max_cand, int* __restrict__ slot_root) {
    const int x = blockIdx.x * 256 + threadIdx.x;
    const int y = blockIdx.y, n = blockIdx.z;
    if (x >= W) return;
    const int rowbase = static_cast<int>((static_cast<long long>(n) * H + y) * W);
    const int p = rowbase + x;
    if (!is_contour_root(label, fg, outer, p)) return;
    const int key = fg[p] ? x : x - 1;  // start pixel column (a hole's start is the pixel left of its first pixel)
    int rank = rowsuf[n * (H + 1) + y];
    for (int q = key + 1; q < W; ++q) {
        if (q == x) continue;
        if (is_contour_root(label, fg, outer, rowbase + q)) {
            const int kq = fg[rowbase + q] ? q : q - 1;
            if (kq > key) ++rank;
        }
    }
    if (rank < max_cand) slot_root[n * kMaxSlots + rank] = y * W + x;
}

// ------------------------------------------------------------------------------------------------ geometry (lane 0)
// 8-neighbour steps of cv2's chain code, s = 0..7: dx = {1, 1, 0, -1, -1, -1, 0, 1}, dy = {0, -1, -1, -1, 0, 1, 1, 1},
// packed two bits per entry (value + 1) so that a step is three ALU ops instead of an indexed constant load whose
// scoreboard wait sat on the border walk's critical path (profiles/r1y_contour_ncu.txt)
__device__ __forceinline__ int c_dx(int s) { return static_cast<int>((0x901Au >> (2 * s)) & 3u) - 1; }
__device__ __forceinline__ int c_dy(int s) { return static_cast<int>((0xA901u >> (2 * s)) & 3u) - 1; }

struct Pt {
    int x, y;
};

// icvFetchContour + CHAIN_APPROX_SIMPLE (see oracle/cv_geom_ref.py trace_border). Returns the vertex count or -1.
// Whole warp: the walk's state is replicated in every lane; the up-to-eight sequential neighbour probes of one step
// (each a dependent global load in the scalar form: 25 % of the kernel's samples, profiles/r1x) are issued by lanes
// 0..7 at once and the first hit in scan order is picked from the ballot.  Lane 0 writes the vertices.
__device__ int trace_border(const uint8_t* fg, int H, int W, int x0, int y0, bool hole, short2* out, int cap, int lane) {
    auto px = [&](int x, int y) -> bool { return x >= 0 && x < W && y >= 0 && y < H && fg[y * W + x] != 0; };
    const int l8 = lane & 7;
    const int s_end = hole ? 0 : 4;
    int s;
    {  // clockwise scan s_end-1, s_end-2, ..., s_end
        const int sl = (s_end - 1 - l8) & 7;
        const unsigned hit = __ballot_sync(0xffffffffu, lane < 8 && px(x0 + c_dx(sl), y0 + c_dy(sl)));
        if (hit == 0u) {
            if (lane == 0) out[0] = make_short2(static_cast<short>(x0), static_cast<short>(y0));
            return 1;
        }
        s = (s_end - 1 - (__ffs(hit) - 1)) & 7;
    }
    const int i1x = x0 + c_dx(s), i1y = y0 + c_dy(s);
    int n = 0;
    int x3 = x0, y3 = y0, ptx = x0, pty = y0, prev_s = s ^ 4;
    for (;;) {
        // counter-clockwise scan s+1, s+2, ...: the pixel we came from is a neighbour, so one of the eight hits
        const int sl = (s + 1 + l8) & 7;
        const unsigned hit = __ballot_sync(0xffffffffu, lane < 8 && px(x3 + c_dx(sl), y3 + c_dy(sl)));
        if (hit == 0u) return -1;  // unreachable on a consistent map; never spin
        s = (s + 1 + (__ffs(hit) - 1)) & 7;
        const int x4 = x3 + c_dx(s), y4 = y3 + c_dy(s);
        if (s != prev_s) {
            if (n >= cap) return -1;
            if (lane == 0) out[n] = make_short2(static_cast<short>(ptx), static_cast<short>(pty));
            ++n;
            prev_s = s;
        }
        ptx += c_dx(s);
        pty += c_dy(s);
        if (x4 == x0 && y4 == y0 && x3 == i1x && y3 == i1y) break;
        x3 = x4;
        y3 = y4;
        s = (s + 4) & 7;
    }
    return n;
}

__device__ __forceinline__ int sgn(long long v) { return (v > 0) - (v < 0); }

// sorted key layout: x (21 bits, biased) | y (21 bits, biased) | original index (22 bits)
__device__ __forceinline__ unsigned long long make_key(int x, int y, int idx) {
    return (static_cast<unsigned long long>(x + 65536) << 43) | (static_cast<unsigned long long>(y + 65536) << 22) |
           static_cast<unsigned long long>(idx);
}
__device__ __forceinline__ int key_x(unsigned long long k) { return static_cast<int>(k >> 43) - 65536; }
__device__ __forceinline__ int key_y(unsigned long long k) { return static_cast<int>((k >> 22) & 0x1FFFFF) - 65536; }
__device__ __forceinline__ int key_i(unsigned long long k) { return static_cast<int>(k & 0x3FFFFF); }

// convhull.cpp Sklansky_ over the sorted keys; stack receives indices into the sorted array
__device__ int sklansky(const unsigned long long* P, int start, int end, int* stack, int nsign, int sign2) {
    const int incr = end > start ? 1 : -1;
    int pprev = start, pcur = pprev + incr, pnext = pcur + incr;
    int stacksize = 3;
    if (start == end || (key_x(P[start]) == key_x(P[end]) && key_y(P[start]) == key_y(P[end]))) {
        stack[0] = start;
        return 1;
    }
    stack[0] = pprev;
    stack[1] = pcur;
    stack[2] = pnext;
    end += incr;
    while (pnext != end) {
        const int cury = key_y(P[pcur]), nexty = key_y(P[pnext]);
        const int by = nexty - cury;
        if (sgn(by) != nsign) {
            const int ax = key_x(P[pcur]) - key_x(P[pprev]);
            const int bx = key_x(P[pnext]) - key_x(P[pcur]);
            const int ay = cury - key_y(P[pprev]);
            const long long convexity = static_cast<long long>(ay) * bx - static_cast<long long>(ax) * by;
            if (sgn(convexity) == sign2 && (ax != 0 || ay != 0)) {
                pprev = pcur;
                pcur = pnext;
                pnext += incr;
                stack[stacksize] = pnext;
                stacksize++;
            } else {
                if (pprev == start) {
                    pcur = pnext;
                    stack[1] = pcur;
                    pnext += incr;
                    stack[2] = pnext;
                } else {
                    stack[stacksize - 2] = pnext;
                    pcur = pprev;
                    pprev = stack[stacksize - 4];
                    stacksize--;
                }
            }
        } else {
            pnext += incr;
            stack[stacksize - 1] = pnext;
        }
    }
    return --stacksize;
}

// cv::convexHull(points, clockwise=false, returnPoints=true) as called by cv::minAreaRect of OpenCV 4.13 (the hull the
// rotating calipers walk is COUNTER-clockwise there: with the clockwise hull of older releases the calipers stop on a
// different but equivalent edge in ~90 % of rotated rectangles and the float32 result differs in its last bits; measured
// against cv2 4.13 in tools/min_area_rect_probe.py).
// P: sorted keys [total]; stack: int[total+2]; hull: int[total] (receives ORIGINAL indices). Returns hull size.
__device__ int convex_hull_ccw(const unsigned long long* P, int total, int* stack, int* hull) {
    int nout = 0;
    int miny = 0, maxy = 0;
    for (int i = 1; i < total; ++i) {
        const int y = key_y(P[i]);
        if (key_y(P[miny]) > y) miny = i;
        if (key_y(P[maxy]) < y) maxy = i;
    }
    if (key_x(P[0]) == key_x(P[total - 1]) && key_y(P[0]) == key_y(P[total - 1])) {
        hull[nout++] = 0;
    } else {
        int* tl = stack;
        int tl_count = sklansky(P, 0, maxy, tl, -1, 1);
        int* tr = stack + tl_count;
        int tr_count = sklansky(P, total - 1, maxy, tr, -1, -1);
        {  // counter-clockwise: swap the two upper chains
            int* t = tl;
            tl = tr;
            tr = t;
            const int c = tl_count;
            tl_count = tr_count;
            tr_count = c;
        }
        for (int i = 0; i < tl_count - 1; ++i) hull[nout++] = tl[i];
        for (int i = tr_count - 1; i > 0; --i) hull[nout++] = tr[i];
        const int stop_idx = tr_count > 2 ? tr[1] : tl_count > 2 ? tl[tl_count - 2] : -1;
        int* bl = stack;
        int bl_count = sklansky(P, 0, miny, bl, 1, -1);
        int* br = stack + bl_count;
        int br_count = sklansky(P, total - 1, miny, br, 1, 1);
        if (stop_idx >= 0) {
            const int check_idx = bl_count > 2 ? bl[1] : (bl_count + br_count > 2 ? br[2 - bl_count] : -1);
            if (check_idx == stop_idx ||
                (check_idx >= 0 && key_x(P[check_idx]) == key_x(P[stop_idx]) && key_y(P[check_idx]) == key_y(P[stop_idx]))) {
                bl_count = bl_count < 2 ? bl_count : 2;
                br_count = br_count < 2 ? br_count : 2;
            }
        }
        for (int i = 0; i < bl_count - 1; ++i) hull[nout++] = bl[i];
        for (int i = br_count - 1; i > 0; --i) hull[nout++] = br[i];
    }
    // sorted position -> original index
    for (int i = 0; i < nout; ++i) hull[i] = key_i(P[hull[i]]);
    if (nout >= 3) {  // cyclic shift that makes the original indices monotone
        int min_idx = 0, max_idx = 0, lt = 0;
        for (int i = 1; i < nout; ++i) {
            const int idx = hull[i];
            lt += hull[i - 1] < idx;
            if (lt > 1 && lt <= i - 2) break;
            if (idx < hull[min_idx]) min_idx = i;
            if (idx > hull[max_idx]) max_idx = i;
        }
        const int mmdist = abs(max_idx - min_idx);
        if ((mmdist == 1 || mmdist == nout - 1) && (lt <= 1 || lt >= nout - 2)) {
            const int ascending = (max_idx + 1) % nout == min_idx;
            const int i0 = ascending ? min_idx : max_idx;
            int j = i0;
            if (i0 > 0) {
                int i;
                for (i = 0; i < nout; ++i) {
                    const int curr = stack[i] = hull[j];
                    const int next_j = j + 1 < nout ? j + 1 : 0;
                    const int next_idx = hull[next_j];
                    if (i < nout - 1 && (ascending != (curr < next_idx))) break;
                    j = next_j;
                }
                if (i == nout)
                    for (int k = 0; k < nout; ++k) hull[k] = stack[k];
            }
        }
    }
    return nout;
}

struct RRect {
    float cx, cy, w, h, ang;
};

// rotcalipers.cpp rotatingCalipers(CALIPERS_MINAREARECT) + cv::minAreaRect tail.  hp[] = hull points (float),
// work = 3*n floats of scratch.
__device__ RRect min_area_rect_hull(const float2* hp, int n, float* work) {
    RRect r{0.f, 0.f, 0.f, 0.f, 0.f};
    double ang = 0.0;
    if (n > 2) {
        float* inv = work;
        float* vx = work + n;
        float* vy = work + 2 * n;
        int left = 0, bottom = 0, right = 0, top = 0;
        float left_x = hp[0].x, right_x = hp[0].x, top_y = hp[0].y, bottom_y = hp[0].y;
        float2 pt0 = hp[0];
        for (int i = 0; i < n; ++i) {
            if (pt0.x < left_x) { left_x = pt0.x; left = i; }
            if (pt0.x > right_x) { right_x = pt0.x; right = i; }
            if (pt0.y > top_y) { top_y = pt0.y; top = i; }
            if (pt0.y < bottom_y) { bottom_y = pt0.y; bottom = i; }
            const float2 pt = hp[i + 1 < n ? i + 1 : 0];
            const double dx = static_cast<double>(pt.x) - static_cast<double>(pt0.x);
            const double dy = static_cast<double>(pt.y) - static_cast<double>(pt0.y);
            vx[i] = static_cast<float>(dx);
            vy[i] = static_cast<float>(dy);
            inv[i] = static_cast<float>(1.0 / sqrt(dx * dx + dy * dy));
            pt0 = pt;
        }
        float orientation = 0.f;
        {
            double ax = vx[n - 1], ay = vy[n - 1];
            for (int i = 0; i < n; ++i) {
                const double bx = vx[i], by = vy[i];
                const double convexity = ax * by - ay * bx;
                if (convexity != 0) {
                    orientation = convexity > 0 ? 1.f : -1.f;
                    break;
                }
                ax = bx;
                ay = by;
            }
        }
        float base_a = orientation, base_b = 0.f;
        int seq[4] = {bottom, right, top, left};
        float minarea = FLT_MAX;
        int b_left = 0, b_bottom = 0;
        float b_a = 0.f, b_w = 0.f, b_b = 0.f, b_h = 0.f;
        for (int k = 0; k < n; ++k) {
            // OpenCV >= 4.5.2 (rotcalipers.cpp firstVecIsRight): the caliper side with the smallest angle to its polygon edge
            // is found from cross-product signs of the edge vectors rotated into a common frame, not by comparing cosines
            float rvx[4], rvy[4];
            rvx[0] = vx[seq[0]];  rvy[0] = vy[seq[0]];
            rvx[1] = vy[seq[1]];  rvy[1] = -vx[seq[1]];
            rvx[2] = -vx[seq[2]]; rvy[2] = -vy[seq[2]];
            rvx[3] = -vy[seq[3]]; rvy[3] = vx[seq[3]];
            int main_element = 0;
            for (int i = 1; i < 4; ++i) {
                const float tx = rvy[i], ty = -rvx[i];  // rotate90CW
                if (__fadd_rn(__fmul_rn(tx, rvx[main_element]), __fmul_rn(ty, rvy[main_element])) < 0.f) main_element = i;
            }
            {
                const int pindex = seq[main_element];
                const float lead_x = __fmul_rn(vx[pindex], inv[pindex]);
                const float lead_y = __fmul_rn(vy[pindex], inv[pindex]);
                switch (main_element) {
                    case 0: base_a = lead_x; base_b = lead_y; break;
                    case 1: base_a = lead_y; base_b = -lead_x; break;
                    case 2: base_a = -lead_x; base_b = -lead_y; break;
                    default: base_a = -lead_y; base_b = lead_x; break;
                }
            }
            seq[main_element] += 1;
            if (seq[main_element] == n) seq[main_element] = 0;
            float dx = __fsub_rn(hp[seq[1]].x, hp[seq[3]].x);
            float dy = __fsub_rn(hp[seq[1]].y, hp[seq[3]].y);
            const float width = __fadd_rn(__fmul_rn(dx, base_a), __fmul_rn(dy, base_b));
            dx = __fsub_rn(hp[seq[2]].x, hp[seq[0]].x);
            dy = __fsub_rn(hp[seq[2]].y, hp[seq[0]].y);
            const float height = __fadd_rn(__fmul_rn(-dx, base_b), __fmul_rn(dy, base_a));
            const float area = __fmul_rn(width, height);
            if (area <= minarea) {
                minarea = area;
                b_left = seq[3];
                b_a = base_a;
                b_w = width;
                b_b = base_b;
                b_h = height;
                b_bottom = seq[0];
            }
        }
        const float A1 = b_a, B1 = b_b, A2 = -b_b, B2 = b_a;
        const float C1 = __fadd_rn(__fmul_rn(A1, hp[b_left].x), __fmul_rn(hp[b_left].y, B1));
        const float C2 = __fadd_rn(__fmul_rn(A2, hp[b_bottom].x), __fmul_rn(hp[b_bottom].y, B2));
        const float idet = __fdiv_rn(1.f, __fsub_rn(__fmul_rn(A1, B2), __fmul_rn(A2, B1)));
        const float ox = __fmul_rn(__fsub_rn(__fmul_rn(C1, B2), __fmul_rn(C2, B1)), idet);
        const float oy = __fmul_rn(__fsub_rn(__fmul_rn(A1, C2), __fmul_rn(A2, C1)), idet);
        const float o1x = __fmul_rn(A1, b_w), o1y = __fmul_rn(B1, b_w);
        const float o2x = __fmul_rn(A2, b_h), o2y = __fmul_rn(B2, b_h);
        r.cx = __fadd_rn(ox, __fmul_rn(__fadd_rn(o1x, o2x), 0.5f));
        r.cy = __fadd_rn(oy, __fmul_rn(__fadd_rn(o1y, o2y), 0.5f));
        r.w = static_cast<float>(sqrt(static_cast<double>(o1x) * o1x + static_cast<double>(o1y) * o1y));
        r.h = static_cast<float>(sqrt(static_cast<double>(o2x) * o2x + static_cast<double>(o2y) * o2y));
        ang = atan2(static_cast<double>(o1y), static_cast<double>(o1x));
    } else if (n == 2) {
        r.cx = __fmul_rn(__fadd_rn(hp[0].x, hp[1].x), 0.5f);
        r.cy = __fmul_rn(__fadd_rn(hp[0].y, hp[1].y), 0.5f);
        const double dx = static_cast<double>(__fsub_rn(hp[1].x, hp[0].x));
        const double dy = static_cast<double>(__fsub_rn(hp[1].y, hp[0].y));
        r.w = static_cast<float>(sqrt(dx * dx + dy * dy));
        ang = atan2(dy, dx);
    } else if (n == 1) {
        r.cx = hp[0].x;
        r.cy = hp[0].y;
    }
    // OpenCV 4.13: the angle is brought into [-90, 0) in DOUBLE by quarter turns, each swapping width and height, and is
    // rounded to float once (observed rule, pinned against cv2 by tests/test_oracle_cpu.py through oracle/cv_geom_ref.py)
    ang = ang * 180 / 3.1415926535897932384626433832795;
    while (ang >= 0.0) {
        ang -= 90.0;
        const float t = r.w;
        r.w = r.h;
        r.h = t;
    }
    while (ang < -90.0) {
        ang += 90.0;
        const float t = r.w;
        r.w = r.h;
        r.h = t;
    }
    r.ang = static_cast<float>(ang);
    return r;
}

// cv::RotatedRect::points
__device__ void box_points(const RRect& r, float2* pt) {
    const double ang = static_cast<double>(r.ang) * 3.1415926535897932384626433832795 / 180.;
    const float b = __fmul_rn(static_cast<float>(cos(ang)), 0.5f);
    const float a = __fmul_rn(static_cast<float>(sin(ang)), 0.5f);
    pt[0].x = __fsub_rn(__fsub_rn(r.cx, __fmul_rn(a, r.h)), __fmul_rn(b, r.w));
    pt[0].y = __fsub_rn(__fadd_rn(r.cy, __fmul_rn(b, r.h)), __fmul_rn(a, r.w));
    pt[1].x = __fsub_rn(__fadd_rn(r.cx, __fmul_rn(a, r.h)), __fmul_rn(b, r.w));
    pt[1].y = __fsub_rn(__fsub_rn(r.cy, __fmul_rn(b, r.h)), __fmul_rn(a, r.w));
    pt[2].x = __fsub_rn(__fmul_rn(2.f, r.cx), pt[0].x);
    pt[2].y = __fsub_rn(__fmul_rn(2.f, r.cy), pt[0].y);
    pt[3].x = __fsub_rn(__fmul_rn(2.f, r.cx), pt[1].x);
    pt[3].y = __fsub_rn(__fmul_rn(2.f, r.cy), pt[1].y);
}

// DBPostProcess.get_mini_boxes ordering (stable sort by x, then pair-wise by y): tl, tr, br, bl
__device__ void mini_box_order(const float2* pt, float2* box) {
    int idx[4] = {0, 1, 2, 3};
    for (int i = 1; i < 4; ++i) {  // insertion sort = stable
        const int v = idx[i];
        int j = i - 1;
        while (j >= 0 && pt[idx[j]].x > pt[v].x) {
            idx[j + 1] = idx[j];
            --j;
        }
        idx[j + 1] = v;
    }
    int i1, i2, i3, i4;
    if (pt[idx[1]].y > pt[idx[0]].y) { i1 = 0; i4 = 1; } else { i1 = 1; i4 = 0; }
    if (pt[idx[3]].y > pt[idx[2]].y) { i2 = 2; i3 = 3; } else { i2 = 3; i3 = 2; }
    box[0] = pt[idx[i1]];
    box[1] = pt[idx[i2]];
    box[2] = pt[idx[i3]];
    box[3] = pt[idx[i4]];
}

// ---- cv2.fillPoly membership for a convex quad with integer vertices (tools/fillpoly_proto.py): pixel on the
// 8-connected Bresenham line of an edge (cv::LineIterator, left-to-right) or inside a 16.16 fixed-point scan-line span
struct QuadFill {
    int vx[4], vy[4];
    long long ex[4], edx[4];  // edge start x (16.16) at y0 and per-row increment
    int ey0[4], ey1[4];
    int ne;
};
__device__ void quad_fill_setup(QuadFill& q) {
    q.ne = 0;
    for (int i = 0; i < 4; ++i) {
        const int j = (i + 3) & 3;  // edge from vertex i-1 to vertex i
        const int x0 = q.vx[j], y0 = q.vy[j], x1 = q.vx[i], y1 = q.vy[i];
        if (y0 == y1) continue;
        const long long X0 = static_cast<long long>(x0) << 16, X1 = static_cast<long long>(x1) << 16;
        const long long d = (X1 - X0) / (y1 - y0);  // C++ truncating division
        const int e = q.ne++;
        q.edx[e] = d;
        if (y0 < y1) { q.ey0[e] = y0; q.ey1[e] = y1; q.ex[e] = X0; } else { q.ey0[e] = y1; q.ey1[e] = y0; q.ex[e] = X1; }
    }
}
// Row form of the line membership (per-pixel definition: oracle/db_post_ref.py, tools/fillpoly_rows.py on_line): the
// pixels of an edge's Bresenham line in row py are one inclusive column interval -- a single pixel for a y-major edge,
// and for an x-major edge the run of columns i with k(i) == kk, i.e. floor(dx (2kk-1) / 2dy) < i <= floor(dx (2kk+1) / 2dy)
// (checked exhaustively against the per-pixel predicate by tools/fillpoly_rows.py).
__device__ __forceinline__ void line_row_interval(int py, int x1, int y1, int x2, int y2, int& lo, int& hi) {
    lo = 1;
    hi = 0;
    int dx = x2 - x1, dy = y2 - y1;
    if (dx < 0) {
        int t = x1; x1 = x2; x2 = t;
        t = y1; y1 = y2; y2 = t;
        dx = -dx;
        dy = -dy;
    }
    int ystep = 1;
    if (dy < 0) {
        dy = -dy;
        ystep = -1;
    }
    const int kk = (py - y1) * ystep;
    if (kk < 0 || kk > dy) return;
    if (dy > dx) {
        const int T = 2 * dx * kk - dy;
        const int k = T <= 0 ? 0 : (T + 2 * dy - 1) / (2 * dy);
        lo = hi = x1 + k;
        return;
    }
    if (dy == 0) {  // horizontal edge (or a single point when dx == 0)
        lo = x1;
        hi = x1 + dx;
        return;
    }
    const int a = kk == 0 ? 0 : (dx * (2 * kk - 1)) / (2 * dy) + 1;
    const int b = min((dx * (2 * kk + 1)) / (2 * dy), dx);
    lo = x1 + a;
    hi = x1 + b;
}
__device__ void quad_fill_row(const QuadFill& q, int py, int* lo, int* hi) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int j = (i + 3) & 3;
        line_row_interval(py, q.vx[j], q.vy[j], q.vx[i], q.vy[i], lo[i], hi[i]);
    }
    long long xa = 0, xb = 0;
    int cnt = 0;
    for (int e = 0; e < q.ne; ++e) {
        if (py >= q.ey0[e] && py < q.ey1[e]) {
            const long long x = q.ex[e] + static_cast<long long>(py - q.ey0[e]) * q.edx[e];
            if (cnt == 0) xa = x; else xb = x;
            ++cnt;
        }
    }
    lo[4] = 1;
    hi[4] = 0;
    if (cnt < 2) return;
    if (xa > xb) {
        const long long t = xa;
        xa = xb;
        xb = t;
    }
    lo[4] = static_cast<int>((xa + 32768) >> 16);
    hi[4] = static_cast<int>(xb >> 16);
}

// ---- Clipper 6.4.2 ClipperOffset (jtRound, etClosedPolygon) of a quad, see oracle/db_post_ref.py
__device__ __forceinline__ long long cround(double v) { return v < 0 ? static_cast<long long>(v - 0.5) : static_cast<long long>(v + 0.5); }

__device__ int clipper_offset_round(const int* qx, const int* qy, int nq, double delta, short2* out, int cap) {
    // AddPath: strip duplicates
    int sx[4], sy[4];
    int hi = nq - 1;
    while (hi > 0 && qx[0] == qx[hi] && qy[0] == qy[hi]) --hi;
    int n = 0;
    sx[n] = qx[0];
    sy[n] = qy[0];
    ++n;
    for (int i = 1; i <= hi; ++i)
        if (sx[n - 1] != qx[i] || sy[n - 1] != qy[i]) {
            sx[n] = qx[i];
            sy[n] = qy[i];
            ++n;
        }
    if (n < 3) return 0;
    {  // FixOrientations
        double a = 0.0;
        int j = n - 1;
        for (int i = 0; i < n; ++i) {
            a += (static_cast<double>(sx[j]) + sx[i]) * (static_cast<double>(sy[j]) - sy[i]);
            j = i;
        }
        if (!(-a * 0.5 >= 0)) {
            for (int i = 0; i < n / 2; ++i) {
                int t = sx[i]; sx[i] = sx[n - 1 - i]; sx[n - 1 - i] = t;
                t = sy[i]; sy[i] = sy[n - 1 - i]; sy[n - 1 - i] = t;
            }
        }
    }
    const double PI = 3.141592653589793238;
    double y;
    const double arc = 0.25;
    if (arc > fabs(delta) * 0.25) y = fabs(delta) * 0.25; else y = arc;
    double steps = PI / acos(1 - y / fabs(delta));
    if (steps > fabs(delta) * PI) steps = fabs(delta) * PI;
    double m_sin = sin(2 * PI / steps);
    const double m_cos = cos(2 * PI / steps);
    const double steps_per_rad = steps / (2 * PI);
    if (delta < 0.0) m_sin = -m_sin;
    double nx[4], ny[4];
    for (int j = 0; j < n; ++j) {
        const int k = (j + 1) % n;
        if (sx[j] == sx[k] && sy[j] == sy[k]) {
            nx[j] = ny[j] = 0;
            continue;
        }
        double dx = static_cast<double>(sx[k] - sx[j]), dy = static_cast<double>(sy[k] - sy[j]);
        const double f = 1.0 / sqrt(dx * dx + dy * dy);
        dx *= f;
        dy *= f;
        nx[j] = dy;
        ny[j] = -dx;
    }
    int m = 0;
    auto push = [&](long long x, long long yv) -> bool {
        if (m >= cap) return false;
        out[m++] = make_short2(static_cast<short>(x), static_cast<short>(yv));
        return true;
    };
    int k = n - 1;
    for (int j = 0; j < n; ++j) {
        double sinA = nx[k] * ny[j] - nx[j] * ny[k];
        bool done = false;
        if (fabs(sinA * delta) < 1.0) {
            const double cosA = nx[k] * nx[j] + ny[j] * ny[k];
            if (cosA > 0) {
                if (!push(cround(sx[j] + nx[k] * delta), cround(sy[j] + ny[k] * delta))) return -1;
                done = true;
            }
        } else if (sinA > 1.0) sinA = 1.0;
        else if (sinA < -1.0) sinA = -1.0;
        if (!done) {
            if (sinA * delta < 0) {
                if (!push(cround(sx[j] + nx[k] * delta), cround(sy[j] + ny[k] * delta))) return -1;
                if (!push(sx[j], sy[j])) return -1;
                if (!push(cround(sx[j] + nx[j] * delta), cround(sy[j] + ny[j] * delta))) return -1;
            } else {
                const double a = atan2(sinA, nx[k] * nx[j] + ny[k] * ny[j]);
                long long st = cround(steps_per_rad * fabs(a));
                if (st < 1) st = 1;
                double X = nx[k], Y = ny[k], X2;
                for (long long i = 0; i < st; ++i) {
                    if (!push(cround(sx[j] + X * delta), cround(sy[j] + Y * delta))) return -1;
                    X2 = X;
                    X = X * m_cos - m_sin * Y;
                    Y = X2 * m_sin + Y * m_cos;
                }
                if (!push(cround(sx[j] + nx[j] * delta), cround(sy[j] + ny[j] * delta))) return -1;
            }
        }
        k = j;
    }
    return m;
}

// warp bitonic sort of n keys (padded to a power of two with ~0) in shared memory
__device__ void warp_sort_keys(unsigned long long* K, int n, int lane) {
    int np2 = 1;
    while (np2 < n) np2 <<= 1;
    for (int i = n + lane; i < np2; i += 32) K[i] = ~0ull;
    __syncwarp();
    for (int k = 2; k <= np2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = lane; i < np2; i += 32) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = K[i], b = K[ixj];
                    const bool up = (i & k) == 0;
                    if ((a > b) == up) {
                        K[i] = b;
                        K[ixj] = a;
                    }
                }
            }
            __syncwarp();
        }
    }
}

// shared-memory plan of one warp (dynamic): keys[kMaxV pow2] | stack[kMaxV+2] | hull[kMaxV] | pts float2[..] ...
struct WarpMem {
    unsigned long long* keys;  // kMaxV
    short2* verts;             // kMaxV  (contour vertices / offset polygon)
    int* stack;                // kMaxV + 2
    int* hull;                 // kMaxV
    float2* hp;                // kMaxV  hull points as float
    float* work;               // 3 * kMaxV
};
constexpr size_t kWarpSmem = kMaxV * 8 + kMaxV * 4 + (kMaxV + 2) * 4 + kMaxV * 4 + kMaxV * 8 + 3 * kMaxV * 4 + 64;

// minAreaRect of the integer points in wm.verts[0..n) (whole warp; result valid in lane 0, broadcast by caller)
__device__ RRect min_area_rect_pts(WarpMem& wm, int n, int lane) {
    for (int i = lane; i < n; i += 32) wm.keys[i] = make_key(wm.verts[i].x, wm.verts[i].y, i);
    __syncwarp();
    warp_sort_keys(wm.keys, n, lane);
    RRect r{0.f, 0.f, 0.f, 0.f, 0.f};
    if (lane == 0) {
        const int hn = convex_hull_ccw(wm.keys, n, wm.stack, wm.hull);
        for (int i = 0; i < hn; ++i) {
            const short2 v = wm.verts[wm.hull[i]];
            wm.hp[i] = make_float2(static_cast<float>(v.x), static_cast<float>(v.y));
        }
        r = min_area_rect_hull(wm.hp, hn, wm.work);
    }
    return r;
}

__global__ void __launch_bounds__(32)
k_db_contour_boxes(const float* __restrict__ prob, int H, int W, const uint8_t* __restrict__ fg, const int* __restrict__ slot_root,
                   const double* __restrict__ src_hw, double box_thresh, double unclip_ratio, int variant, float* __restrict__ slot_box,
                   int* __restrict__ slot_valid, int* __restrict__ overflow) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int slot = blockIdx.x, n = blockIdx.y, lane = threadIdx.x;
    const int root = slot_root[n * kMaxSlots + slot];
    if (lane == 0) slot_valid[n * kMaxSlots + slot] = 0;
    if (root < 0) return;
    WarpMem wm;
    {
        unsigned char* p = smem_raw;
        wm.keys = reinterpret_cast<unsigned long long*>(p); p += kMaxV * 8;
        wm.hp = reinterpret_cast<float2*>(p); p += kMaxV * 8;
        wm.verts = reinterpret_cast<short2*>(p); p += kMaxV * 4;
        wm.stack = reinterpret_cast<int*>(p); p += (kMaxV + 2) * 4;
        wm.hull = reinterpret_cast<int*>(p); p += kMaxV * 4;
        wm.work = reinterpret_cast<float*>(p);
    }
    const uint8_t* fgp = fg + static_cast<long long>(n) * H * W;
    const float* pp = prob + static_cast<long long>(n) * H * W;
    // ---- 1. trace the border
    int nv;
    {
        const int ry = root / W, rx = root - ry * W;
        const bool hole = fgp[root] == 0;
        nv = trace_border(fgp, H, W, hole ? rx - 1 : rx, ry, hole, wm.verts, kMaxV, lane);
        if (nv < 0 && lane == 0) atomicAdd(overflow, 1);
    }
    if (nv <= 0) return;
    __syncwarp();
    // ---- 2. first minAreaRect -> mini box
    RRect r = min_area_rect_pts(wm, nv, lane);
    float2 box[4];
    int keep = 0;
    if (lane == 0) {
        float2 pt[4];
        box_points(r, pt);
        mini_box_order(pt, box);
        const float sside = fminf(r.w, r.h);
        keep = !(sside < 3.f);
    }
    keep = __shfl_sync(0xffffffffu, keep, 0);
    if (!keep) return;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        box[i].x = __shfl_sync(0xffffffffu, box[i].x, 0);
        box[i].y = __shfl_sync(0xffffffffu, box[i].y, 0);
    }
    // ---- 3. box_score_fast: mean of prob over the fillPoly mask of the truncated quad inside its bounding box
    float mnx = fminf(fminf(box[0].x, box[1].x), fminf(box[2].x, box[3].x));
    float mxx = fmaxf(fmaxf(box[0].x, box[1].x), fmaxf(box[2].x, box[3].x));
    float mny = fminf(fminf(box[0].y, box[1].y), fminf(box[2].y, box[3].y));
    float mxy = fmaxf(fmaxf(box[0].y, box[1].y), fmaxf(box[2].y, box[3].y));
    auto clampi = [](long long v, int lo, int hi) -> int { return static_cast<int>(v < lo ? lo : (v > hi ? hi : v)); };
    const int xmin = clampi(static_cast<long long>(floorf(mnx)), 0, W - 1), xmax = clampi(static_cast<long long>(ceilf(mxx)), 0, W - 1);
    const int ymin = clampi(static_cast<long long>(floorf(mny)), 0, H - 1), ymax = clampi(static_cast<long long>(ceilf(mxy)), 0, H - 1);
    QuadFill q;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        q.vx[i] = static_cast<int>(__fsub_rn(box[i].x, static_cast<float>(xmin)));  // astype(int32): truncation
        q.vy[i] = static_cast<int>(__fsub_rn(box[i].y, static_cast<float>(ymin)));
    }
    quad_fill_setup(q);
    const int bw = xmax - xmin + 1, bh = ymax - ymin + 1;
    // Row by row: the mask's membership in a row is the union of five inclusive column intervals (the four edge lines
    // and the scan-line span), computed once per row; the per-pixel work is then a few compares and one load.
    // (ncu, profiles/r1x: the per-pixel form of this test was 55 % of this kernel's samples.)
    {  // pull the box's rows of the probability map towards L1 first: the row loop below otherwise pays one full
       // memory latency per row (21 % of the kernel's samples after the row-interval rewrite)
        const int lpr = (bw * 4 + 127) / 128 + 1;  // 128-byte lines per row (rows are not line-aligned)
        for (int i = lane; i < bh * lpr; i += 32) {
            const int py = i / lpr, l = i - py * lpr;
            const float* a = pp + (ymin + py) * W + min(xmin + l * 32, xmax);
            asm volatile("prefetch.global.L1 [%0];" ::"l"(a));
        }
    }
    double sum = 0.0;
    int cnt = 0;
    for (int py = 0; py < bh; ++py) {
        int lo[5], hi[5];
        quad_fill_row(q, py, lo, hi);
        int mn = bw, mx = -1;
#pragma unroll
        for (int e = 0; e < 5; ++e)
            if (lo[e] <= hi[e]) {
                mn = min(mn, lo[e]);
                mx = max(mx, hi[e]);
            }
        mn = max(mn, 0);
        mx = min(mx, bw - 1);
        const float* prow = pp + (ymin + py) * W + xmin;
        for (int px = mn + lane; px <= mx; px += 64) {
            const int p1 = px + 32;
            bool in0 = false, in1 = false;
#pragma unroll
            for (int e = 0; e < 5; ++e) {
                in0 |= px >= lo[e] && px <= hi[e];
                in1 |= p1 >= lo[e] && p1 <= hi[e];
            }
            in1 &= p1 <= mx;
            const float v0 = in0 ? __ldg(prow + px) : 0.f;  // both loads in flight before either is summed
            const float v1 = in1 ? __ldg(prow + p1) : 0.f;
            sum += static_cast<double>(v0);
            sum += static_cast<double>(v1);
            cnt += static_cast<int>(in0) + static_cast<int>(in1);
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, o);
        cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
    }
    const double score = cnt > 0 ? sum / cnt : 0.0;
    if (box_thresh > score) return;
    // ---- 4. unclip (lane 0) -> offset polygon in wm.verts
    int no = 0;
    if (lane == 0) {
        double area = 0.0, length = 0.0;
        for (int i = 0; i < 4; ++i) {
            const int j = (i + 1) & 3;
            const double xi = box[i].x, yi = box[i].y, xj = box[j].x, yj = box[j].y;
            area += xi * yj - yi * xj;
            length += sqrt((xi - xj) * (xi - xj) + (yi - yj) * (yi - yj));
        }
        area = fabs(area) * 0.5;
        const double distance = area * unclip_ratio / length;
        int qx[4], qy[4];
        for (int i = 0; i < 4; ++i) {
            qx[i] = static_cast<int>(box[i].x);  // pyclipper <cInt> cast: truncation toward zero
            qy[i] = static_cast<int>(box[i].y);
        }
        no = clipper_offset_round(qx, qy, 4, distance, wm.verts, kMaxOff);
        if (no < 0) atomicAdd(overflow, 1);
    }
    no = __shfl_sync(0xffffffffu, no, 0);
    if (no <= 0) return;
    __syncwarp();
    // ---- 5. second minAreaRect, scaling, ordering, filters (lane 0)
    r = min_area_rect_pts(wm, no, lane);
    if (lane != 0) return;
    float2 pt[4];
    box_points(r, pt);
    mini_box_order(pt, box);
    if (fminf(r.w, r.h) < 5.f) return;
    const double src_h = src_hw[n * 2], src_w = src_hw[n * 2 + 1];
    if (variant == 1) {
        // the in-tree DBNet back-end (db_net/ocr_detection_utils.py:190-205): box.astype(np.int32) FIRST (truncation), then
        // np.clip(np.round(box / width * dest), 0, dest) in float64 stored back into the int32 box; corner order of
        // get_mini_boxes is kept and filter_tag_det_res does not exist there
        float* ob = slot_box + (static_cast<long long>(n) * kMaxSlots + slot) * 8;
        for (int i = 0; i < 4; ++i) {
            const double xi = static_cast<double>(static_cast<int>(box[i].x)), yi = static_cast<double>(static_cast<int>(box[i].y));
            const double vx = fmin(fmax(rint(__dmul_rn(__ddiv_rn(xi, static_cast<double>(W)), src_w)), 0.0), src_w);
            const double vy = fmin(fmax(rint(__dmul_rn(__ddiv_rn(yi, static_cast<double>(H)), src_h)), 0.0), src_h);
            ob[2 * i] = static_cast<float>(static_cast<int>(vx));
            ob[2 * i + 1] = static_cast<float>(static_cast<int>(vy));
        }
        slot_valid[n * kMaxSlots + slot] = 1;
        return;
    }
    float bx[4], by[4];
    for (int i = 0; i < 4; ++i) {
        // np.clip(np.round(box / width * dest_width), 0, dest_width): float32 division, float64 product, rint
        double vx = static_cast<double>(__fdiv_rn(box[i].x, static_cast<float>(W))) * src_w;
        double vy = static_cast<double>(__fdiv_rn(box[i].y, static_cast<float>(H))) * src_h;
        vx = fmin(fmax(rint(vx), 0.0), src_w);
        vy = fmin(fmax(rint(vy), 0.0), src_h);
        // stored into the float32 box, then .astype(np.int16)
        bx[i] = static_cast<float>(static_cast<short>(static_cast<int>(static_cast<float>(vx))));
        by[i] = static_cast<float>(static_cast<short>(static_cast<int>(static_cast<float>(vy))));
    }
    // order_points_clockwise: stable argsort by x; left pair by y -> (tl, bl); right pair by y -> (tr, br)
    int idx[4] = {0, 1, 2, 3};
    for (int i = 1; i < 4; ++i) {
        const int v = idx[i];
        int j = i - 1;
        while (j >= 0 && bx[idx[j]] > bx[v]) {
            idx[j + 1] = idx[j];
            --j;
        }
        idx[j + 1] = v;
    }
    int tl = idx[0], bl = idx[1], tr = idx[2], br = idx[3];
    if (by[tl] > by[bl]) { const int t = tl; tl = bl; bl = t; }
    if (by[tr] > by[br]) { const int t = tr; tr = br; br = t; }
    const int ord[4] = {tl, tr, br, bl};
    float ox[4], oy[4];
    const int img_h = static_cast<int>(src_h), img_w = static_cast<int>(src_w);
    for (int i = 0; i < 4; ++i) {
        ox[i] = static_cast<float>(static_cast<int>(fminf(fmaxf(bx[ord[i]], 0.f), static_cast<float>(img_w - 1))));
        oy[i] = static_cast<float>(static_cast<int>(fminf(fmaxf(by[ord[i]], 0.f), static_cast<float>(img_h - 1))));
    }
    auto norm_i = [](float dx, float dy) -> int { return static_cast<int>(sqrtf(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)))); };
    const int rect_w = norm_i(ox[0] - ox[1], oy[0] - oy[1]);
    const int rect_h = norm_i(ox[0] - ox[3], oy[0] - oy[3]);
    if (rect_w <= 3 || rect_h <= 3) return;
    float* ob = slot_box + (static_cast<long long>(n) * kMaxSlots + slot) * 8;
    for (int i = 0; i < 4; ++i) {
        ob[2 * i] = ox[i];
        ob[2 * i + 1] = oy[i];
    }
    slot_valid[n * kMaxSlots + slot] = 1;
}

// left-pack the valid boxes of each page in slot order (= the reference's contour order)
__global__ void __launch_bounds__(32)
k_db_compact(const float* __restrict__ slot_box, const int* __restrict__ slot_valid, int max_out, float* __restrict__ boxes_out,
             int* __restrict__ counts_out) {
    const int n = blockIdx.x, lane = threadIdx.x;
    int count = 0;
    for (int s0 = 0; s0 < kMaxSlots; s0 += 32) {
        const int s = s0 + lane;
        const bool v = s < max_out && slot_valid[n * kMaxSlots + s] != 0;  // slots >= max_out hold stale flags
        const unsigned m = __ballot_sync(0xffffffffu, v);
        if (v) {
            const int pos = count + __popc(m & ((1u << lane) - 1u));
            if (pos < max_out) {
                const float4* src = reinterpret_cast<const float4*>(slot_box + (static_cast<long long>(n) * kMaxSlots + s) * 8);
                float4* dst = reinterpret_cast<float4*>(boxes_out + (static_cast<long long>(n) * max_out + pos) * 8);
                dst[0] = src[0];
                dst[1] = src[1];
            }
        }
        count += __popc(m);
    }
    if (lane == 0) counts_out[n] = count < max_out ? count : max_out;
}

int ensure_ws(Engine* e, DbWs* ws, int N, int H, int W) {
    if (ws->N == N && ws->H == H && ws->W == W) return 0;
    for (void* p : ws->mem) cudaFree(p);
    ws->mem.clear();
    auto alloc = [&](void** p, size_t bytes) -> int {
        cudaError_t st = cudaMalloc(p, bytes);
        if (st != cudaSuccess) return set_err(e, DV_ERR_CUDA, "db_boxes: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(st));
        ws->mem.push_back(*p);
        return 0;
    };
    const size_t npx = static_cast<size_t>(N) * H * W;
    DV_TRY(alloc(reinterpret_cast<void**>(&ws->label), npx * 4));
    DV_TRY(alloc(reinterpret_cast<void**>(&ws->fg), npx));
    DV_TRY(alloc(reinterpret_cast<void**>(&ws->outer), npx));
    DV_TRY(alloc(reinterpret_cast<void**>(&ws->rowcnt), static_cast<size_t>(N) * (H + 1) * 4));
    DV_TRY(alloc(reinterpret_cast<void**>(&ws->rowsuf), static_cast<size_t>(N) * (H + 1) * 4));
    DV_TRY(alloc(reinterpret_cast<void**>(&ws->ncont), static_cast<size_t>(N) * 4));
    DV_TRY(alloc(reinterpret_cast<void**>(&ws->slot_root), static_cast<size_t>(N) * kMaxSlots * 4));
    DV_TRY(alloc(reinterpret_cast<void**>(&ws->slot_box), static_cast<size_t>(N) * kMaxSlots * 8 * 4));
    DV_TRY(alloc(reinterpret_cast<void**>(&ws->slot_valid), static_cast<size_t>(N) * kMaxSlots * 4));
    DV_TRY(alloc(reinterpret_cast<void**>(&ws->src_hw), static_cast<size_t>(N) * 2 * 8));
    DV_TRY(alloc(reinterpret_cast<void**>(&ws->overflow), 4));
    ws->N = N;
    ws->H = H;
    ws->W = W;
    return 0;
}

}  // namespace

int db_boxes(Engine* e, const float* prob, int N, int H, int W, const double* src_hw_host, float thresh, double box_thresh,
             double unclip_ratio, int max_candidates, float* boxes_out, int32_t* counts_out, int32_t* overflow_host, int variant) {
    if (N == 0) return 0;
    if (!prob || !src_hw_host || !boxes_out || !counts_out || N < 0 || H <= 0 || W <= 0)
        return set_err(e, DV_ERR_ARG, "db_boxes: bad arguments");
    if (max_candidates <= 0 || max_candidates > kMaxSlots)
        return set_err(e, DV_ERR_UNSUPPORTED, "db_boxes: max_candidates must be in 1..%d", kMaxSlots);
    if (H > 32000 || W > 32000 || static_cast<long long>(N) * H * W > 0x7fffffffLL)
        return set_err(e, DV_ERR_UNSUPPORTED, "db_boxes: map too large");
    auto it = e->aux.find("db_post");
    if (it == e->aux.end()) it = e->aux.emplace("db_post", std::unique_ptr<Model>(new DbWs())).first;
    DbWs* ws = static_cast<DbWs*>(it->second.get());
    DV_TRY(ensure_ws(e, ws, N, H, W));
    static DeviceOnce attr_once;
    if (attr_once.need(e->device)) {
        DV_CUDA(e, cudaFuncSetAttribute(k_db_contour_boxes, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(kWarpSmem)));
        attr_once.mark(e->device);
    }
    cudaStream_t s = e->stream;
    const size_t npx = static_cast<size_t>(N) * H * W;
    DV_CUDA(e, cudaMemcpyAsync(ws->src_hw, src_hw_host, static_cast<size_t>(N) * 2 * 8, cudaMemcpyHostToDevice, s));
    DV_CUDA(e, cudaMemsetAsync(ws->outer, 0, npx, s));
    DV_CUDA(e, cudaMemsetAsync(ws->rowcnt, 0, static_cast<size_t>(N) * (H + 1) * 4, s));
    DV_CUDA(e, cudaMemsetAsync(ws->ncont, 0, static_cast<size_t>(N) * 4, s));
    DV_CUDA(e, cudaMemsetAsync(ws->slot_root, 0xff, static_cast<size_t>(N) * kMaxSlots * 4, s));
    DV_CUDA(e, cudaMemsetAsync(ws->overflow, 0, 4, s));
    const dim3 grid((W + 255) / 256, H, N);
    const double px = static_cast<double>(npx);
    e->launch_begin("k_db_label_init", "db_post", 0.0, px * (4 + 4 + 1));
    k_db_label_init<<<grid, 256, 0, s>>>(prob, H, W, thresh, ws->label, ws->fg);
    e->launch_end();
    e->launch_begin("k_db_label_merge", "db_post", 0.0, px * (4 + 1));
    k_db_label_merge<<<grid, 256, 0, s>>>(H, W, ws->label, ws->fg);
    e->launch_end();
    e->launch_begin("k_db_label_flatten", "db_post", 0.0, px * (4 + 4 + 1));
    k_db_label_flatten<<<grid, 256, 0, s>>>(H, W, ws->label, ws->fg, ws->outer);
    e->launch_end();
    e->launch_begin("k_db_enumerate", "db_post", 0.0, px * 4);
    k_db_enumerate<<<grid, 256, 0, s>>>(H, W, ws->label, ws->fg, ws->outer, ws->rowcnt, ws->ncont);
    e->launch_end();
    e->launch_begin("k_db_rowsuffix", "db_post", 0.0, static_cast<double>(N) * H * 8);
    k_db_rowsuffix<<<N, 32, 0, s>>>(H, ws->rowcnt, ws->rowsuf);
    e->launch_end();
    e->launch_begin("k_db_rank", "db_post", 0.0, px * 4);
    k_db_rank<<<grid, 256, 0, s>>>(H, W, ws->label, ws->fg, ws->outer, ws->rowsuf, max_candidates, ws->slot_root);
    e->launch_end();
    e->launch_begin("k_db_contour_boxes", "db_post", 0.0, px * 1.0);
    k_db_contour_boxes<<<dim3(max_candidates, N), 32, kWarpSmem, s>>>(prob, H, W, ws->fg, ws->slot_root, ws->src_hw, box_thresh,
                                                                      unclip_ratio, variant, ws->slot_box, ws->slot_valid, ws->overflow);
    e->launch_end();
    e->launch_begin("k_db_compact", "db_post", 0.0, static_cast<double>(N) * kMaxSlots * 36);
    k_db_compact<<<N, 32, 0, s>>>(ws->slot_box, ws->slot_valid, max_candidates, boxes_out, counts_out);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    if (overflow_host) {
        DV_CUDA(e, cudaMemcpyAsync(overflow_host, ws->overflow, 4, cudaMemcpyDeviceToHost, s));
        DV_CUDA(e, cudaStreamSynchronize(s));
    }
    return 0;
}

}  // namespace dv
