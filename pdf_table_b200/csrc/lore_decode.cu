// centernet_heatmap_decode: Lore head maps -> sorted table cells, entirely on the GPU.
//
// Follows the reference process_detect_output (lore/lineless_table_process.py:592-655): corner_decode :97-124,
// ctdet_4ps_decode :127-267 (incl. the sequential `wiz_rev` corner snapping :178-236), _nms :66-73, _topk :76-94,
// is_group_faster_faster :355-379, find4ps :329-338, dist :341-345, cc_match :240-251, _get_4ps_feat's clamp :55-61,
// the score re-sort :255-262, ctdet_4ps_post_process / affine_transform :489-507, 387-390, filter :568-582 and
// normalized_ps :585-589.  The CPU mirror is oracle/lore_decode_ref.py (bit-exact against the reference on the
// golden cases); this file must stay bit-exact against that mirror: float32 steps use explicit _rn intrinsics so
// nothing is contracted into an FMA the reference does not perform.
//
// GPU formulation
//   1. k_lore_peaks    : one thread per (pixel, class): 3x3 max test (-inf border) fused with the score gate; every
//                        survivor appends a 64-bit key (score bits | ~index).  The reference's top-K over all H*W
//                        positions only ever *uses* rows above the 0.2 / 0.3 gates, so the gate is applied first
//                        and the dense `heat * keep` map is never written.
//   2. k_lore_sort     : one CTA per (image, class): bitonic sort of the keys in shared memory (descending score,
//                        ascending index on ties), truncation to K / MK, gather of reg / wh / st at the peaks.
//   3. k_lore_wiz_rev  : one warp per cell.  Lanes test 32 corner boxes at a time (bounding-box overlap, then the
//                        strict point-in-quad predicate in float64); the matching corners are then applied in corner
//                        order, because the reference's update rule is order dependent.
//   4. k_lore_finalize : one CTA per image: re-sort by the penalised score, inverse affine to source pixels, integer
//                        position features, and the gather indices of the logical-feature step (`ax` follows the
//                        re-sorted order, the four `cr` corner indices do NOT -- reference quirk :255-262, :644).
//   5. k_lore_gather_logi : logi_feat[j] = ax[:, ax_idx[j]] + sum_k cr[:, cr_idx[j][k]] from dense maps.
#include <math.h>

#include "engine.h"

namespace dv {

namespace {

constexpr int kCap = 16384;       // peak candidates per (image, class): > H*W/4 of a 256x256 map
constexpr int kSortThreads = 1024;

struct LoreWs : Model {
    int N = 0, K = 0, MK = 0;
    unsigned long long* keys = nullptr;  // [N][2][kCap]
    int* nkeys = nullptr;                // [N][2]
    // cells (class 0), K per image
    int* cell_n = nullptr;       // [N]
    float* cell_score = nullptr;  // [N][K]   (penalised in place by wiz_rev)
    int* cell_idx = nullptr;     // [N][K]
    float* cell_box = nullptr;   // [N][K][8]  original corners
    float* cell_rev = nullptr;   // [N][K][8]  snapped corners
    int* cell_cc = nullptr;      // [N][K][4]
    // corners (class 1), MK per image, structure of arrays for coalesced lane access
    int* cor_n = nullptr;        // [N]
    float* cor_score = nullptr;  // [N][MK]
    float* cor_xy = nullptr;     // [N][2][MK]
    float* cor_box = nullptr;    // [N][8][MK]
    double* trans = nullptr;     // [N][6]
    int* overflow = nullptr;
    std::vector<void*> mem;
    ~LoreWs() override {
        for (void* p : mem) cudaFree(p);
    }
};

// Strided view of the four small head maps: element (n, c, pixel) = base[n*img + c*chan + pixel*pix]
struct MapView {
    const float* p;
    long long img, chan, pix;
    __device__ __forceinline__ float at(int n, int c, int pixel) const { return p[n * img + c * chan + pixel * pix]; }
};

__global__ void __launch_bounds__(256)
k_lore_peaks(MapView hm, int H, int W, float gate_cell, float gate_corner, unsigned long long* __restrict__ keys,
             int* __restrict__ nkeys, int* __restrict__ overflow) {
    const int pix = blockIdx.x * 256 + threadIdx.x;
    const int cls = blockIdx.y, n = blockIdx.z;
    if (pix >= H * W) return;
    const float v = hm.at(n, cls, pix);
    if (!(v >= (cls == 0 ? gate_cell : gate_corner))) return;
    const int y = pix / W, x = pix - y * W;
    for (int dy = -1; dy <= 1; ++dy) {
        const int yy = y + dy;
        if (yy < 0 || yy >= H) continue;
        for (int dx = -1; dx <= 1; ++dx) {
            const int xx = x + dx;
            if (xx < 0 || xx >= W) continue;
            if (hm.at(n, cls, yy * W + xx) > v) return;  // not the 3x3 maximum (ties survive, as `hmax == heat`)
        }
    }
    const int slot = atomicAdd(&nkeys[n * 2 + cls], 1);
    if (slot >= kCap) {
        atomicExch(overflow, 1);
        return;
    }
    keys[(static_cast<size_t>(n) * 2 + cls) * kCap + slot] =
        (static_cast<unsigned long long>(__float_as_uint(v)) << 32) | (0xffffffffu - static_cast<unsigned>(pix));
}

// in-CTA bitonic sort, descending, of `n2` (power of two) keys in shared memory
__device__ void bitonic_desc(unsigned long long* s, int n2) {
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < n2; i += blockDim.x) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const unsigned long long a = s[i], b = s[ixj];
                    const bool desc = (i & k) == 0;
                    if (desc ? (a < b) : (a > b)) {
                        s[i] = b;
                        s[ixj] = a;
                    }
                }
            }
            __syncthreads();
        }
    }
}

__global__ void __launch_bounds__(kSortThreads)
k_lore_sort(MapView reg, MapView wh, MapView st, int W, int K, int MK, const unsigned long long* __restrict__ keys,
            const int* __restrict__ nkeys, int* __restrict__ cell_n, float* __restrict__ cell_score,
            int* __restrict__ cell_idx, float* __restrict__ cell_box, float* __restrict__ cell_rev, int* __restrict__ cor_n,
            float* __restrict__ cor_score, float* __restrict__ cor_xy, float* __restrict__ cor_box) {
    extern __shared__ unsigned long long s_keys[];
    const int cls = blockIdx.x, n = blockIdx.y;
    int cnt = nkeys[n * 2 + cls];
    if (cnt > kCap) cnt = kCap;
    int n2 = 1;
    while (n2 < cnt) n2 <<= 1;
    const unsigned long long* src = keys + (static_cast<size_t>(n) * 2 + cls) * kCap;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) s_keys[i] = i < cnt ? src[i] : 0ull;
    __syncthreads();
    bitonic_desc(s_keys, n2);
    const int keep = min(cnt, cls == 0 ? K : MK);
    if (threadIdx.x == 0) (cls == 0 ? cell_n : cor_n)[n] = keep;
    for (int i = threadIdx.x; i < keep; i += blockDim.x) {
        const unsigned long long key = s_keys[i];
        const float score = __uint_as_float(static_cast<unsigned>(key >> 32));
        const int pix = static_cast<int>(0xffffffffu - static_cast<unsigned>(key & 0xffffffffu));
        const int y = pix / W, x = pix - y * W;
        const float xs = __fadd_rn(static_cast<float>(x), reg.at(n, 0, pix));
        const float ys = __fadd_rn(static_cast<float>(y), reg.at(n, 1, pix));
        if (cls == 0) {
            const size_t o = static_cast<size_t>(n) * K + i;
            cell_score[o] = score;
            cell_idx[o] = pix;
            for (int k = 0; k < 8; ++k) {
                const float b = __fsub_rn((k & 1) ? ys : xs, wh.at(n, k, pix));
                cell_box[o * 8 + k] = b;
                cell_rev[o * 8 + k] = b;
            }
        } else {
            cor_score[static_cast<size_t>(n) * MK + i] = score;
            cor_xy[(static_cast<size_t>(n) * 2 + 0) * MK + i] = xs;
            cor_xy[(static_cast<size_t>(n) * 2 + 1) * MK + i] = ys;
            for (int k = 0; k < 8; ++k)
                cor_box[(static_cast<size_t>(n) * 8 + k) * MK + i] = __fsub_rn((k & 1) ? ys : xs, st.at(n, k, pix));
        }
    }
}

// oracle/lore_decode_ref.py point_strictly_in_polygon: crossing number in float64, boundary points are outside
__device__ bool point_strictly_in_quad(double px, double py, const double* qx, const double* qy) {
    bool inside = false;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
        const double x1 = qx[a], y1 = qy[a], x2 = qx[(a + 1) & 3], y2 = qy[(a + 1) & 3];
        const double cross = __dsub_rn(__dmul_rn(x2 - x1, py - y1), __dmul_rn(y2 - y1, px - x1));
        if (cross == 0.0 && fmin(x1, x2) <= px && px <= fmax(x1, x2) && fmin(y1, y2) <= py && py <= fmax(y1, y2)) return false;
        if ((y1 > py) != (y2 > py)) {
            const double t = __dsub_rn(__dmul_rn(px - x1, y2 - y1), __dmul_rn(x2 - x1, py - y1));
            if ((t < 0) == ((y2 - y1) > 0)) inside = !inside;
        }
    }
    return inside;
}

__device__ __forceinline__ float dist2(float x1, float y1, float x2, float y2) {
    const float dx = __fsub_rn(x1, x2), dy = __fsub_rn(y1, y2);
    return __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
}

__global__ void __launch_bounds__(128)
k_lore_wiz_rev(int K, int MK, const int* __restrict__ cell_n, float* __restrict__ cell_score, const float* __restrict__ cell_box,
               float* __restrict__ cell_rev, const int* __restrict__ cor_n, const float* __restrict__ cor_score,
               const float* __restrict__ cor_xy, const float* __restrict__ cor_box) {
    const int n = blockIdx.y;
    const int lane = threadIdx.x & 31;
    const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (i >= cell_n[n]) return;
    const size_t o = static_cast<size_t>(n) * K + i;
    const float score = cell_score[o];
    if (!(score >= 0.2f)) return;  // the reference breaks out of the (sorted) cell loop here
    float b[8], r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = b[k] = cell_box[o * 8 + k];
    double qx[4], qy[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        qx[k] = b[2 * k];
        qy[k] = b[2 * k + 1];
    }
    const float bxmin = fminf(fminf(b[0], b[2]), fminf(b[4], b[6])), bxmax = fmaxf(fmaxf(b[0], b[2]), fmaxf(b[4], b[6]));
    const float bymin = fminf(fminf(b[1], b[3]), fminf(b[5], b[7])), bymax = fmaxf(fmaxf(b[1], b[3]), fmaxf(b[5], b[7]));
    const int nc = cor_n[n];
    const float* cs = cor_score + static_cast<size_t>(n) * MK;
    const float* cxy = cor_xy + static_cast<size_t>(n) * 2 * MK;
    const float* cb = cor_box + static_cast<size_t>(n) * 8 * MK;
    int count = 0;
    for (int j0 = 0; j0 < nc; j0 += 32) {
        const int j = j0 + lane;
        bool live = j < nc && cs[j] >= 0.3f;  // sorted descending: the first failing corner ends the reference's loop
        const unsigned live_mask = __ballot_sync(0xffffffffu, live);
        bool hit = false;
        float cx = 0.f, cy = 0.f;
        if (live) {
            float g[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) g[k] = cb[static_cast<size_t>(k) * MK + j];
            const float gxmin = fminf(fminf(g[0], g[2]), fminf(g[4], g[6])), gxmax = fmaxf(fmaxf(g[0], g[2]), fmaxf(g[4], g[6]));
            const float gymin = fminf(fminf(g[1], g[3]), fminf(g[5], g[7])), gymax = fmaxf(fmaxf(g[1], g[3]), fmaxf(g[5], g[7]));
            if (!(bxmin > gxmax || gxmin > bxmax || bymin > gymax || gymin > bymax)) {
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (!hit && point_strictly_in_quad(g[2 * k], g[2 * k + 1], qx, qy)) hit = true;
            }
            cx = cxy[j];
            cy = cxy[MK + j];
        }
        unsigned hits = __ballot_sync(0xffffffffu, hit);
        while (hits) {  // apply in ascending corner order; every lane keeps the same state
            const int src = __ffs(hits) - 1;
            hits &= hits - 1;
            const float px = __shfl_sync(0xffffffffu, cx, src), py = __shfl_sync(0xffffffffu, cy, src);
            int q = 0;  // find4ps: first minimum of the squared distance to the ORIGINAL corners
            float best = dist2(b[0], b[1], px, py);
#pragma unroll
            for (int k = 1; k < 4; ++k) {
                const float d = dist2(b[2 * k], b[2 * k + 1], px, py);
                if (d < best) {
                    best = d;
                    q = k;
                }
            }
            float ox = b[0], oy = b[1], rx = r[0], ry = r[1];
#pragma unroll
            for (int k = 1; k < 4; ++k)
                if (q == k) {
                    ox = b[2 * k];
                    oy = b[2 * k + 1];
                    rx = r[2 * k];
                    ry = r[2 * k + 1];
                }
            bool take;
            if (rx == ox && ry == oy) take = true;
            else take = dist2(ox, oy, rx, ry) >= dist2(ox, oy, px, py);
            if (take) {
                ++count;
#pragma unroll
                for (int k = 0; k < 4; ++k)
                    if (q == k) {
                        r[2 * k] = px;
                        r[2 * k + 1] = py;
                    }
            }
        }
        if (live_mask != 0xffffffffu) break;
    }
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < 8; ++k) cell_rev[o * 8 + k] = r[k];
        if (count <= 2) cell_score[o] = __fmul_rn(score, 0.4f);
    }
}

__global__ void __launch_bounds__(kSortThreads)
k_lore_finalize(int H, int W, int K, int wiz_rev, float vis_thresh, int batch_clamp, const int* __restrict__ cell_n,
                const float* __restrict__ cell_score, const int* __restrict__ cell_idx, const float* __restrict__ cell_rev,
                const double* __restrict__ trans, float* __restrict__ polygons, float* __restrict__ scores_out,
                int32_t* __restrict__ dets_feat, int32_t* __restrict__ ax_idx, int32_t* __restrict__ cr_idx,
                int32_t* __restrict__ counts, int32_t* __restrict__ rows) {
    extern __shared__ unsigned long long s_keys[];
    __shared__ int s_valid;
    const int n = blockIdx.x;
    const int cnt = cell_n[n];
    int n2 = 1;
    while (n2 < cnt) n2 <<= 1;
    if (threadIdx.x == 0) s_valid = 0;
    for (int i = threadIdx.x; i < n2; i += blockDim.x)
        s_keys[i] = i < cnt ? (static_cast<unsigned long long>(__float_as_uint(cell_score[static_cast<size_t>(n) * K + i])) << 32) |
                                  (0xffffffffu - static_cast<unsigned>(i))
                            : 0ull;
    __syncthreads();
    if (wiz_rev) bitonic_desc(s_keys, n2);  // without wiz_rev the reference keeps the top-K order
    const double t0 = trans[n * 6 + 0], t1 = trans[n * 6 + 1], t2 = trans[n * 6 + 2];
    const double t3 = trans[n * 6 + 3], t4 = trans[n * 6 + 4], t5 = trans[n * 6 + 5];
    int valid = 0;
    for (int j = threadIdx.x; j < cnt; j += blockDim.x) {
        const unsigned long long key = s_keys[j];
        const int src = static_cast<int>(0xffffffffu - static_cast<unsigned>(key & 0xffffffffu));
        const float score = __uint_as_float(static_cast<unsigned>(key >> 32));
        const size_t o = static_cast<size_t>(n) * K + j;
        const size_t os = static_cast<size_t>(n) * K + src;
        scores_out[o] = score;
        if (score >= vis_thresh) ++valid;
        ax_idx[o] = cell_idx[os];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float x = cell_rev[os * 8 + 2 * k], y = cell_rev[os * 8 + 2 * k + 1];
            const double dx = x, dy = y;
            polygons[o * 8 + 2 * k] = static_cast<float>(__dadd_rn(__dadd_rn(__dmul_rn(t0, dx), __dmul_rn(t1, dy)), t2));
            polygons[o * 8 + 2 * k + 1] = static_cast<float>(__dadd_rn(__dadd_rn(__dmul_rn(t3, dx), __dmul_rn(t4, dy)), t5));
            dets_feat[o * 8 + 2 * k] = min(max(static_cast<int>(x), 0), 255);  // int32 truncation, then [0, 255]
            dets_feat[o * 8 + 2 * k + 1] = min(max(static_cast<int>(y), 0), 255);
            // cc_match of the UN-sorted row j (reference quirk): round(x + W * round(y)), float32, half-to-even
            const float ux = cell_rev[o * 8 + 2 * k], uy = cell_rev[o * 8 + 2 * k + 1];
            const float m = rintf(__fadd_rn(ux, __fmul_rn(static_cast<float>(W), rintf(uy))));
            long long c = static_cast<long long>(m);
            if (!(c < static_cast<long long>(H) * W)) c = batch_clamp;
            if (c < 0) c = 0;
            cr_idx[o * 4 + k] = static_cast<int32_t>(c);
        }
    }
    if (valid) atomicAdd(&s_valid, valid);
    __syncthreads();
    if (threadIdx.x == 0) {
        counts[n] = s_valid;
        if (rows) rows[n] = cnt;
    }
}

__global__ void __launch_bounds__(256)
k_lore_gather_logi(const float* __restrict__ ax, const float* __restrict__ cr, int C, int HW, int K, const int32_t* __restrict__ counts,
                   const int32_t* __restrict__ ax_idx, const int32_t* __restrict__ cr_idx, float* __restrict__ out) {
    const int n = blockIdx.y, j = blockIdx.x;
    if (j >= counts[n]) return;
    const size_t o = static_cast<size_t>(n) * K + j;
    const int a = ax_idx[o];
    const int c0 = cr_idx[o * 4], c1 = cr_idx[o * 4 + 1], c2 = cr_idx[o * 4 + 2], c3 = cr_idx[o * 4 + 3];
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const float* axc = ax + (static_cast<size_t>(n) * C + c) * HW;
        const float* crc = cr + (static_cast<size_t>(n) * C + c) * HW;
        float s = __fadd_rn(0.f, crc[c0]);
        s = __fadd_rn(s, crc[c1]);
        s = __fadd_rn(s, crc[c2]);
        s = __fadd_rn(s, crc[c3]);
        out[o * C + c] = __fadd_rn(axc[a], s);
    }
}

// ------------------------------------------------------------------------------------------------ CenterNet
// OCRTableCenterNetPostProcessor (center_net/processer_centernet.py:170-204) after bbox_decode / gbox_decode (the peaks and
// sort kernels above): inverse affine of cells (4 corners) and vertices (position + 4 predicted cell centres) to source
// pixels, group_bbox_by_gbox (center_net/table_process.py:278-333), score > 0.3 filter, sort by 0.01 * mean_x + mean_y.
__device__ __forceinline__ float affine_x(const double* t, float x, float y) {
    return static_cast<float>(__dadd_rn(__dadd_rn(__dmul_rn(t[0], static_cast<double>(x)), __dmul_rn(t[1], static_cast<double>(y))), t[2]));
}
__device__ __forceinline__ float affine_y(const double* t, float x, float y) {
    return static_cast<float>(__dadd_rn(__dadd_rn(__dmul_rn(t[3], static_cast<double>(x)), __dmul_rn(t[4], static_cast<double>(y))), t[5]));
}

__global__ void __launch_bounds__(256)
k_cn_transform(int K, int MK, const int* __restrict__ cell_n, float* __restrict__ cell_box, const int* __restrict__ cor_n,
               float* __restrict__ cor_xy, float* __restrict__ cor_box, const double* __restrict__ trans) {
    const int n = blockIdx.y;
    const double* t = trans + n * 6;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < cell_n[n]) {
        float* b = cell_box + (static_cast<size_t>(n) * K + i) * 8;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float x = b[2 * k], y = b[2 * k + 1];
            b[2 * k] = affine_x(t, x, y);
            b[2 * k + 1] = affine_y(t, x, y);
        }
    }
    if (i < cor_n[n]) {
        float* xy = cor_xy + static_cast<size_t>(n) * 2 * MK;
        float* cb = cor_box + static_cast<size_t>(n) * 8 * MK;
        const float x = xy[i], y = xy[MK + i];
        xy[i] = affine_x(t, x, y);
        xy[MK + i] = affine_y(t, x, y);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const float cx = cb[static_cast<size_t>(2 * k) * MK + i], cy = cb[static_cast<size_t>(2 * k + 1) * MK + i];
            cb[static_cast<size_t>(2 * k) * MK + i] = affine_x(t, cx, cy);
            cb[static_cast<size_t>(2 * k + 1) * MK + i] = affine_y(t, cx, cy);
        }
    }
}

__device__ __forceinline__ float cross_f32(float ax, float ay, float bx, float by) {  // ax * by - ay * bx, numpy float32 steps
    return __fsub_rn(__fmul_rn(ax, by), __fmul_rn(ay, bx));
}

// One warp per cell.  A corner j of cell k snaps to the FIRST (vertex, centre) pair in the reference's loop order whose
// centre lies strictly inside the (original) cell, whose vertex is >= 2 px from that centre, for which j is the cell corner
// nearest to the vertex, and that is closer than half the cell size.  (`sum(sign[k]) == 4: continue` never changes a result.)
__global__ void __launch_bounds__(128)
k_cn_group(int K, int MK, const int* __restrict__ cell_n, const float* __restrict__ cell_box, float* __restrict__ cell_rev,
           const int* __restrict__ cor_n, const float* __restrict__ cor_xy, const float* __restrict__ cor_box) {
    const int n = blockIdx.y, lane = threadIdx.x & 31;
    const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (k >= cell_n[n]) return;
    const size_t o = static_cast<size_t>(n) * K + k;
    float b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = cell_box[o * 8 + i];
    const float w = __fmul_rn(__fadd_rn(fabsf(__fsub_rn(b[6], b[0])), fabsf(__fsub_rn(b[4], b[2]))), 0.5f);
    const float h = __fmul_rn(__fadd_rn(fabsf(__fsub_rn(b[3], b[1])), fabsf(__fsub_rn(b[5], b[7]))), 0.5f);
    const double lim = static_cast<double>(__fmul_rn(0.5f, fmaxf(w, h)));
    const float* vxy = cor_xy + static_cast<size_t>(n) * 2 * MK;
    const float* vb = cor_box + static_cast<size_t>(n) * 8 * MK;
    const int nv = cor_n[n];
    int best[4] = {0x7fffffff, 0x7fffffff, 0x7fffffff, 0x7fffffff};
    for (int v = lane; v < nv; v += 32) {
        const float vx = vxy[v], vy = vxy[MK + v];
        float dc[4];  // squared distances (float32 steps) from the vertex to the cell corners
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const float dx = __fsub_rn(vx, b[2 * j]), dy = __fsub_rn(vy, b[2 * j + 1]);
            dc[j] = __fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy));
        }
        int jm = 0;
#pragma unroll
        for (int j = 1; j < 4; ++j)
            if (dc[j] < dc[jm]) jm = j;
        float dmin = dc[0];
#pragma unroll
        for (int j = 1; j < 4; ++j)
            if (jm == j) dmin = dc[j];
        if (!(sqrt(static_cast<double>(dmin)) < lim)) continue;
        for (int i = 0; i < 4; ++i) {
            const float cx = vb[static_cast<size_t>(2 * i) * MK + v], cy = vb[static_cast<size_t>(2 * i + 1) * MK + v];
            const float dx = __fsub_rn(vx, cx), dy = __fsub_rn(vy, cy);
            if (__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)) < 4.0f) continue;  // get_distance(vertex, centre) < 2
            const float a0 = cross_f32(__fsub_rn(b[2], b[0]), __fsub_rn(b[3], b[1]), __fsub_rn(cx, b[0]), __fsub_rn(cy, b[1]));
            const float a1 = cross_f32(__fsub_rn(b[4], b[2]), __fsub_rn(b[5], b[3]), __fsub_rn(cx, b[2]), __fsub_rn(cy, b[3]));
            const float a2 = cross_f32(__fsub_rn(b[6], b[4]), __fsub_rn(b[7], b[5]), __fsub_rn(cx, b[4]), __fsub_rn(cy, b[5]));
            const float a3 = cross_f32(__fsub_rn(b[0], b[6]), __fsub_rn(b[1], b[7]), __fsub_rn(cx, b[6]), __fsub_rn(cy, b[7]));
            const bool in = (a0 > 0 && a1 > 0 && a2 > 0 && a3 > 0) || (a0 < 0 && a1 < 0 && a2 < 0 && a3 < 0);
            if (!in) continue;
            const int key = v * 4 + i;
#pragma unroll
            for (int j = 0; j < 4; ++j)
                if (jm == j && key < best[j]) best[j] = key;
        }
    }
#pragma unroll
    for (int j = 0; j < 4; ++j)
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) best[j] = min(best[j], __shfl_xor_sync(0xffffffffu, best[j], off));
    if (lane == 0) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float x = b[2 * j], y = b[2 * j + 1];
            if (best[j] != 0x7fffffff) {
                x = vxy[best[j] >> 2];
                y = vxy[MK + (best[j] >> 2)];
            }
            cell_rev[o * 8 + 2 * j] = x;
            cell_rev[o * 8 + 2 * j + 1] = y;
        }
    }
}

__global__ void __launch_bounds__(kSortThreads)
k_cn_finalize(int K, float score_thr, const int* __restrict__ cell_n, const float* __restrict__ cell_score, const float* __restrict__ cell_rev,
              float* __restrict__ polygons, int32_t* __restrict__ counts) {
    extern __shared__ unsigned long long s_keys[];
    __shared__ int s_cnt;
    const int n = blockIdx.x;
    const int cnt = cell_n[n];
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    int n2 = 1;
    while (n2 < cnt) n2 <<= 1;
    for (int i = threadIdx.x; i < n2; i += blockDim.x) {
        unsigned long long key = 0ull;  // sorts last (descending sort of the inverted key)
        if (i < cnt && cell_score[static_cast<size_t>(n) * K + i] > score_thr) {
            const float* b = cell_rev + (static_cast<size_t>(n) * K + i) * 8;
            // key = 0.01 * (sum(x) / 4) + sum(y) / 4 with Python's left-to-right float32 sums
            const float sx = __fadd_rn(__fadd_rn(__fadd_rn(b[0], b[2]), b[4]), b[6]);
            const float sy = __fadd_rn(__fadd_rn(__fadd_rn(b[1], b[3]), b[5]), b[7]);
            const float kf = __fadd_rn(__fmul_rn(0.01f, __fmul_rn(sx, 0.25f)), __fmul_rn(sy, 0.25f));
            unsigned u = __float_as_uint(kf);
            u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);  // order-preserving map of float32 to unsigned
            key = (static_cast<unsigned long long>(~u) << 32) | (0xffffffffu - static_cast<unsigned>(i));  // ascending key, then ascending i
            atomicAdd(&s_cnt, 1);
        }
        s_keys[i] = key;
    }
    __syncthreads();
    bitonic_desc(s_keys, n2);
    const int keep = s_cnt;
    for (int j = threadIdx.x; j < keep; j += blockDim.x) {
        const int src = static_cast<int>(0xffffffffu - static_cast<unsigned>(s_keys[j] & 0xffffffffu));
        const float4* b = reinterpret_cast<const float4*>(cell_rev + (static_cast<size_t>(n) * K + src) * 8);
        float4* dst = reinterpret_cast<float4*>(polygons + (static_cast<size_t>(n) * K + j) * 8);
        dst[0] = b[0];
        dst[1] = b[1];
    }
    if (threadIdx.x == 0) counts[n] = keep;
}

int ensure_ws(Engine* e, LoreWs* ws, int N, int K, int MK) {
    if (ws->N >= N && ws->K == K && ws->MK == MK) return 0;
    for (void* p : ws->mem) cudaFree(p);
    ws->mem.clear();
    auto alloc = [&](void** p, size_t bytes) -> int {
        cudaError_t st = cudaMalloc(p, bytes);
        if (st != cudaSuccess) return set_err(e, DV_ERR_CUDA, "lore_decode: cudaMalloc(%zu): %s", bytes, cudaGetErrorString(st));
        ws->mem.push_back(*p);
        return 0;
    };
    const size_t n = static_cast<size_t>(N);
    DV_TRY(alloc(reinterpret_cast<void**>(&ws->keys), n * 2 * kCap * 8));
    DV_TRY(alloc(reinterpret_cast<void**>(&ws->nkeys), n * 2 * 4));
    DV_TRY(alloc(reinterpret_cast<void**>(&ws->cell_n), n * 4));
    DV_TRY(alloc(reinterpret_cast<void**>(&ws->cell_score), n * K * 4));
    DV_TRY(alloc(reinterpret_cast<void**>(&ws->cell_idx), n * K * 4));
    DV_TRY(alloc(reinterpret_cast<void**>(&ws->cell_box), n * K * 32));
    DV_TRY(alloc(reinterpret_cast<void**>(&ws->cell_rev), n * K * 32));
    DV_TRY(alloc(reinterpret_cast<void**>(&ws->cor_n), n * 4));
    DV_TRY(alloc(reinterpret_cast<void**>(&ws->cor_score), n * MK * 4));
    DV_TRY(alloc(reinterpret_cast<void**>(&ws->cor_xy), n * MK * 8));
    DV_TRY(alloc(reinterpret_cast<void**>(&ws->cor_box), n * MK * 32));
    DV_TRY(alloc(reinterpret_cast<void**>(&ws->trans), n * 48));
    DV_TRY(alloc(reinterpret_cast<void**>(&ws->overflow), 4));
    ws->N = N;
    ws->K = K;
    ws->MK = MK;
    return 0;
}

}  // namespace

int lore_decode(Engine* e, const LoreMaps& maps, int N, int H, int W, const double* trans_host, int K, int MK, int wiz_rev,
                float vis_thresh, float* polygons, float* scores, int32_t* dets_feat, int32_t* ax_idx, int32_t* cr_idx,
                int32_t* counts, int32_t* rows, int32_t* overflow_host) {
    if (N == 0) return 0;
    if (!maps.hm || !maps.reg || !maps.wh || !maps.st || !trans_host || !polygons || !scores || !dets_feat || !ax_idx ||
        !cr_idx || !counts || N < 0 || H <= 0 || W <= 0)
        return set_err(e, DV_ERR_ARG, "lore_decode: bad arguments");
    if (K <= 0 || K > 4096 || MK <= 0 || MK > kCap) return set_err(e, DV_ERR_UNSUPPORTED, "lore_decode: K in 1..4096, MK in 1..%d", kCap);
    if (static_cast<long long>(H) * W > (1 << 24)) return set_err(e, DV_ERR_UNSUPPORTED, "lore_decode: map too large");
    auto it = e->aux.find("lore_decode");
    if (it == e->aux.end()) it = e->aux.emplace("lore_decode", std::unique_ptr<Model>(new LoreWs())).first;
    LoreWs* ws = static_cast<LoreWs*>(it->second.get());
    DV_TRY(ensure_ws(e, ws, N, K, MK));
    static DeviceOnce attr_once;
    if (attr_once.need(e->device)) {
        DV_CUDA(e, cudaFuncSetAttribute(k_lore_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, kCap * 8));
        attr_once.mark(e->device);
    }
    cudaStream_t s = e->stream;
    DV_CUDA(e, cudaMemcpyAsync(ws->trans, trans_host, static_cast<size_t>(N) * 48, cudaMemcpyHostToDevice, s));
    DV_CUDA(e, cudaMemsetAsync(ws->nkeys, 0, static_cast<size_t>(N) * 8, s));
    DV_CUDA(e, cudaMemsetAsync(ws->overflow, 0, 4, s));
    const MapView hm{maps.hm, maps.img_stride[0], maps.chan_stride[0], maps.pix_stride[0]};
    const MapView reg{maps.reg, maps.img_stride[1], maps.chan_stride[1], maps.pix_stride[1]};
    const MapView wh{maps.wh, maps.img_stride[2], maps.chan_stride[2], maps.pix_stride[2]};
    const MapView st{maps.st, maps.img_stride[3], maps.chan_stride[3], maps.pix_stride[3]};
    // the x0.4 penalty only lowers scores, so a cell below min(0.2, vis_thresh) can never be selected
    const float gate_cell = wiz_rev ? fminf(0.2f, vis_thresh) : vis_thresh;
    const double px = static_cast<double>(N) * H * W;
    e->launch_begin("k_lore_peaks", "lore_decode", 0.0, px * 2 * 4);
    k_lore_peaks<<<dim3((H * W + 255) / 256, 2, N), 256, 0, s>>>(hm, H, W, gate_cell, 0.3f, ws->keys, ws->nkeys, ws->overflow);
    e->launch_end();
    e->launch_begin("k_lore_sort", "lore_decode", 0.0, static_cast<double>(N) * (K + MK) * 20 * 4);
    k_lore_sort<<<dim3(2, N), kSortThreads, kCap * 8, s>>>(reg, wh, st, W, K, MK, ws->keys, ws->nkeys, ws->cell_n, ws->cell_score,
                                                           ws->cell_idx, ws->cell_box, ws->cell_rev, ws->cor_n, ws->cor_score,
                                                           ws->cor_xy, ws->cor_box);
    e->launch_end();
    if (wiz_rev) {
        e->launch_begin("k_lore_wiz_rev", "lore_decode", 0.0, static_cast<double>(N) * (K * 72.0 + MK * 44.0));
        k_lore_wiz_rev<<<dim3((K + 3) / 4, N), 128, 0, s>>>(K, MK, ws->cell_n, ws->cell_score, ws->cell_box, ws->cell_rev, ws->cor_n,
                                                            ws->cor_score, ws->cor_xy, ws->cor_box);
        e->launch_end();
    }
    e->launch_begin("k_lore_finalize", "lore_decode", 0.0, static_cast<double>(N) * K * 120.0);
    k_lore_finalize<<<N, kSortThreads, 4096 * 8, s>>>(H, W, K, wiz_rev, vis_thresh, /*batch_clamp=*/0, ws->cell_n, ws->cell_score,
                                                      ws->cell_idx, ws->cell_rev, ws->trans, polygons, scores, dets_feat, ax_idx,
                                                      cr_idx, counts, rows);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    if (overflow_host) {
        DV_CUDA(e, cudaMemcpyAsync(overflow_host, ws->overflow, 4, cudaMemcpyDeviceToHost, s));
        DV_CUDA(e, cudaStreamSynchronize(s));
    }
    return 0;
}

int lore_gather_logi(Engine* e, const float* ax, const float* cr, int N, int C, int H, int W, int K, const int32_t* counts,
                     const int32_t* ax_idx, const int32_t* cr_idx, float* logi_feat) {
    if (N == 0) return 0;
    if (!ax || !cr || !counts || !ax_idx || !cr_idx || !logi_feat) return set_err(e, DV_ERR_ARG, "lore_gather_logi: null argument");
    e->launch_begin("k_lore_gather_logi", "lore_decode", 0.0, static_cast<double>(N) * K * C * 6 * 4);
    k_lore_gather_logi<<<dim3(K, N), 256, 0, e->stream>>>(ax, cr, C, H * W, K, counts, ax_idx, cr_idx, logi_feat);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

int centernet_decode(Engine* e, const LoreMaps& maps, int N, int H, int W, const double* trans_host, int K, int MK, float score_thr,
                     float* polygons, int32_t* counts, int32_t* overflow_host) {
    if (N == 0) return 0;
    if (!maps.hm || !maps.reg || !maps.wh || !maps.st || !trans_host || !polygons || !counts || N < 0 || H <= 0 || W <= 0)
        return set_err(e, DV_ERR_ARG, "centernet_decode: bad arguments");
    if (K <= 0 || K > 4096 || MK <= 0 || MK > kCap) return set_err(e, DV_ERR_UNSUPPORTED, "centernet_decode: K in 1..4096, MK in 1..%d", kCap);
    auto it = e->aux.find("lore_decode");
    if (it == e->aux.end()) it = e->aux.emplace("lore_decode", std::unique_ptr<Model>(new LoreWs())).first;
    LoreWs* ws = static_cast<LoreWs*>(it->second.get());
    DV_TRY(ensure_ws(e, ws, N, K, MK));
    static DeviceOnce attr_once;
    if (attr_once.need(e->device)) {
        DV_CUDA(e, cudaFuncSetAttribute(k_lore_sort, cudaFuncAttributeMaxDynamicSharedMemorySize, kCap * 8));
        attr_once.mark(e->device);
    }
    cudaStream_t s = e->stream;
    DV_CUDA(e, cudaMemcpyAsync(ws->trans, trans_host, static_cast<size_t>(N) * 48, cudaMemcpyHostToDevice, s));
    DV_CUDA(e, cudaMemsetAsync(ws->nkeys, 0, static_cast<size_t>(N) * 8, s));
    DV_CUDA(e, cudaMemsetAsync(ws->overflow, 0, 4, s));
    const MapView hm{maps.hm, maps.img_stride[0], maps.chan_stride[0], maps.pix_stride[0]};
    const MapView reg{maps.reg, maps.img_stride[1], maps.chan_stride[1], maps.pix_stride[1]};
    const MapView wh{maps.wh, maps.img_stride[2], maps.chan_stride[2], maps.pix_stride[2]};
    const MapView st{maps.st, maps.img_stride[3], maps.chan_stride[3], maps.pix_stride[3]};
    const double px = static_cast<double>(N) * H * W;
    e->launch_begin("k_lore_peaks", "centernet_decode", 0.0, px * 2 * 4);
    k_lore_peaks<<<dim3((H * W + 255) / 256, 2, N), 256, 0, s>>>(hm, H, W, score_thr, score_thr, ws->keys, ws->nkeys, ws->overflow);
    e->launch_end();
    e->launch_begin("k_lore_sort", "centernet_decode", 0.0, static_cast<double>(N) * (K + MK) * 20 * 4);
    k_lore_sort<<<dim3(2, N), kSortThreads, kCap * 8, s>>>(reg, wh, st, W, K, MK, ws->keys, ws->nkeys, ws->cell_n, ws->cell_score, ws->cell_idx,
                                                           ws->cell_box, ws->cell_rev, ws->cor_n, ws->cor_score, ws->cor_xy, ws->cor_box);
    e->launch_end();
    e->launch_begin("k_cn_transform", "centernet_decode", 0.0, static_cast<double>(N) * (K * 64.0 + MK * 80.0));
    k_cn_transform<<<dim3((max(K, MK) + 255) / 256, N), 256, 0, s>>>(K, MK, ws->cell_n, ws->cell_box, ws->cor_n, ws->cor_xy, ws->cor_box, ws->trans);
    e->launch_end();
    e->launch_begin("k_cn_group", "centernet_decode", 0.0, static_cast<double>(N) * (K * 64.0 + MK * 40.0));
    k_cn_group<<<dim3((K + 3) / 4, N), 128, 0, s>>>(K, MK, ws->cell_n, ws->cell_box, ws->cell_rev, ws->cor_n, ws->cor_xy, ws->cor_box);
    e->launch_end();
    e->launch_begin("k_cn_finalize", "centernet_decode", 0.0, static_cast<double>(N) * K * 72.0);
    k_cn_finalize<<<N, kSortThreads, 4096 * 8, s>>>(K, score_thr, ws->cell_n, ws->cell_score, ws->cell_rev, polygons, counts);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    if (overflow_host) {
        DV_CUDA(e, cudaMemcpyAsync(overflow_host, ws->overflow, 4, cudaMemcpyDeviceToHost, s));
        DV_CUDA(e, cudaStreamSynchronize(s));
    }
    return 0;
}

}  // namespace dv
