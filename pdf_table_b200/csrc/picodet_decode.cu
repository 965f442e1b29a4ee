// picodet_anchor_decode: PicoDet head outputs -> layout boxes, entirely on the GPU.
//
// Follows the reference OCRPicodetPostProcessor.__call__ (picodet/processor_picodet.py:184-298): anchor centres
// :207-214, DFL softmax-integral :216-221, per-level top-k by the best class score :223-228, box decode :231, per-class
// score threshold + hard_nms :240-256 (hard_nms :301-331: 200 best candidates, greedy IoU suppression with eps 1e-5,
// keep_top_k), warp_boxes :136-158 (clip against the ORIGINAL image size, one float32 round trip) and the division by
// the scale factor :266-272.  dtypes as in the reference: class scores and the softmax are float32, everything after
// `softmax * arange` is float64.  The CPU mirror is oracle/picodet_ref.py (bit-exact against the reference class).
//
// GPU formulation
//   1. k_pico_select : one CTA per (level, image).  Only anchors whose best class score passes the threshold can ever
//      reach the NMS, so the gate comes first; if more than nms_top_k anchors pass, an in-CTA bitonic sort keeps the
//      best nms_top_k (the reference's argsort over all 7600 anchors of a level is never materialised).  The selected
//      anchors' boxes are decoded in float64.
//   2. k_pico_nms    : one CTA per (class, image): gather the class's candidates, sort by score, keep 200, greedy NMS
//      with the 200 x 200 IoU tests spread over the threads, emit at most keep_top_k boxes.
//   3. k_pico_pack   : concatenate the classes in ascending order per image (the reference's output order).
#include <math.h>

#include "engine.h"

namespace dv {

namespace {

constexpr int kLevels = 4;
constexpr int kMaxAnchors = 16384;  // per level (800x608 / 8 -> 7600)
constexpr int kTopK = 1000;         // nms_top_k capacity per level
constexpr int kNmsCand = 200;       // hard_nms candidate_size
constexpr int kMaxKeep = 100;       // keep_top_k capacity
constexpr int kMaxClasses = 16;

struct PicoWs : Model {
    int N = 0;
    double* cand_box = nullptr;  // [N][kLevels*kTopK][4]
    int* cand_anchor = nullptr;  // [N][kLevels*kTopK]  (level << 24 | anchor)
    int* cand_n = nullptr;       // [N][kLevels]
    double* cls_rows = nullptr;  // [N][kMaxClasses][kMaxKeep][6]
    int* cls_n = nullptr;        // [N][kMaxClasses]
    float* meta = nullptr;       // [N][4] org_h, org_w, ratio_h, ratio_w
    std::vector<void*> mem;
    ~PicoWs() override {
        for (void* p : mem) cudaFree(p);
    }
};

struct LevelPtrs {
    const float* score[kLevels];
    const float* dfl[kLevels];
    int stride[kLevels];
    int fw[kLevels], hw[kLevels];
};

__device__ void bitonic_desc64(unsigned long long* s, int n2) {
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

// DFL integral of one side: float32 softmax (numpy's 8-wide pairwise sum order), float64 expectation
__device__ double dfl_distance(const float* __restrict__ logits, int bins) {
    float mx = logits[0];
    for (int k = 1; k < bins; ++k) mx = fmaxf(mx, logits[k]);
    float e[16];
    for (int k = 0; k < bins; ++k) e[k] = expf(__fsub_rn(logits[k], mx));
    float sum;
    if (bins == 8) {
        sum = __fadd_rn(__fadd_rn(__fadd_rn(e[0], e[1]), __fadd_rn(e[2], e[3])), __fadd_rn(__fadd_rn(e[4], e[5]), __fadd_rn(e[6], e[7])));
    } else {
        sum = e[0];
        for (int k = 1; k < bins; ++k) sum = __fadd_rn(sum, e[k]);
    }
    double t[16];
    for (int k = 0; k < bins; ++k) t[k] = static_cast<double>(__fdiv_rn(e[k], sum)) * static_cast<double>(k);
    if (bins == 8) return __dadd_rn(__dadd_rn(__dadd_rn(t[0], t[1]), __dadd_rn(t[2], t[3])), __dadd_rn(__dadd_rn(t[4], t[5]), __dadd_rn(t[6], t[7])));
    double s = t[0];
    for (int k = 1; k < bins; ++k) s = __dadd_rn(s, t[k]);
    return s;
}

__global__ void __launch_bounds__(256)
k_pico_select(LevelPtrs lp, int C, int bins, float score_thr, int top_k, double* __restrict__ cand_box, int* __restrict__ cand_anchor,
              int* __restrict__ cand_n) {
    extern __shared__ unsigned long long s_keys[];
    __shared__ int s_cnt;
    const int lvl = blockIdx.x, n = blockIdx.y;
    const int hw = lp.hw[lvl];
    const float* sc = lp.score[lvl] + static_cast<size_t>(n) * hw * C;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    for (int a = threadIdx.x; a < hw; a += blockDim.x) {
        float m = sc[static_cast<size_t>(a) * C];
        for (int c = 1; c < C; ++c) m = fmaxf(m, sc[static_cast<size_t>(a) * C + c]);
        if (m > score_thr) {
            const int slot = atomicAdd(&s_cnt, 1);
            s_keys[slot] = (static_cast<unsigned long long>(__float_as_uint(m)) << 32) | (0xffffffffu - static_cast<unsigned>(a));
        }
    }
    __syncthreads();
    const int cnt = s_cnt;
    int n2 = 1;
    while (n2 < cnt) n2 <<= 1;
    for (int i = cnt + threadIdx.x; i < n2; i += blockDim.x) s_keys[i] = 0ull;
    __syncthreads();
    bitonic_desc64(s_keys, n2);
    const int keep = min(cnt, top_k);
    if (threadIdx.x == 0) cand_n[n * kLevels + lvl] = keep;
    const int stride = lp.stride[lvl], fw = lp.fw[lvl];
    const float* dfl = lp.dfl[lvl] + static_cast<size_t>(n) * hw * 4 * bins;
    for (int i = threadIdx.x; i < keep; i += blockDim.x) {
        const int a = static_cast<int>(0xffffffffu - static_cast<unsigned>(s_keys[i] & 0xffffffffu));
        const int y = a / fw, x = a - y * fw;
        const double cx = (static_cast<double>(x) + 0.5) * stride, cy = (static_cast<double>(y) + 0.5) * stride;
        const float* d = dfl + static_cast<size_t>(a) * 4 * bins;
        const size_t o = (static_cast<size_t>(n) * kLevels + lvl) * kTopK + i;
        cand_box[o * 4 + 0] = cx - dfl_distance(d, bins) * stride;
        cand_box[o * 4 + 1] = cy - dfl_distance(d + bins, bins) * stride;
        cand_box[o * 4 + 2] = cx + dfl_distance(d + 2 * bins, bins) * stride;
        cand_box[o * 4 + 3] = cy + dfl_distance(d + 3 * bins, bins) * stride;
        cand_anchor[o] = (lvl << 24) | a;
    }
}

__device__ __forceinline__ double area_of(double x0, double y0, double x1, double y1) {
    return fmax(x1 - x0, 0.0) * fmax(y1 - y0, 0.0);
}

__global__ void __launch_bounds__(256)
k_pico_nms(LevelPtrs lp, int C, float score_thr, double iou_thr, int keep_top_k, const double* __restrict__ cand_box,
           const int* __restrict__ cand_anchor, const int* __restrict__ cand_n, const float* __restrict__ meta,
           double* __restrict__ cls_rows, int* __restrict__ cls_n) {
    __shared__ unsigned long long s_keys[4096];  // class candidates (gated rows of the 4 x nms_top_k selection)
    __shared__ double s_box[kNmsCand][4];
    __shared__ float s_score[kNmsCand];
    __shared__ unsigned char s_dead[kNmsCand];
    __shared__ int s_cnt;
    const int c = blockIdx.x, n = blockIdx.y;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    // candidates of this class, in the reference's concatenation order (level, then descending best score)
    for (int lvl = 0; lvl < kLevels; ++lvl) {
        const int cn = cand_n[n * kLevels + lvl];
        const float* sc = lp.score[lvl] + static_cast<size_t>(n) * lp.hw[lvl] * C;
        for (int i = threadIdx.x; i < cn; i += blockDim.x) {
            const int slot = lvl * kTopK + i;
            const int a = cand_anchor[static_cast<size_t>(n) * kLevels * kTopK + slot] & 0xffffff;
            const float p = sc[static_cast<size_t>(a) * C + c];
            if (p > score_thr) {
                const int k = atomicAdd(&s_cnt, 1);
                if (k < 4096) s_keys[k] = (static_cast<unsigned long long>(__float_as_uint(p)) << 32) | static_cast<unsigned>(slot);
            }
        }
    }
    __syncthreads();
    const int cnt = min(s_cnt, 4096);
    int n2 = 1;
    while (n2 < cnt) n2 <<= 1;
    for (int i = cnt + threadIdx.x; i < n2; i += blockDim.x) s_keys[i] = 0ull;
    __syncthreads();
    bitonic_desc64(s_keys, n2);
    const int m = min(cnt, kNmsCand);
    for (int i = threadIdx.x; i < m; i += blockDim.x) {
        const int slot = static_cast<int>(s_keys[i] & 0xffffffffu);
        const double* b = cand_box + (static_cast<size_t>(n) * kLevels * kTopK + slot) * 4;
        s_box[i][0] = b[0];
        s_box[i][1] = b[1];
        s_box[i][2] = b[2];
        s_box[i][3] = b[3];
        s_score[i] = __uint_as_float(static_cast<unsigned>(s_keys[i] >> 32));
        s_dead[i] = 0;
    }
    __syncthreads();
    int picked = 0;
    const float org_h = meta[n * 4 + 0], org_w = meta[n * 4 + 1], ratio_h = meta[n * 4 + 2], ratio_w = meta[n * 4 + 3];
    for (int i = 0; i < m && picked < keep_top_k; ++i) {
        if (s_dead[i]) continue;  // uniform: shared state, read after the barrier below
        const double x0 = s_box[i][0], y0 = s_box[i][1], x1 = s_box[i][2], y1 = s_box[i][3];
        if (threadIdx.x == 0) {
            double* r = cls_rows + ((static_cast<size_t>(n) * kMaxClasses + c) * kMaxKeep + picked) * 6;
            r[0] = static_cast<double>(c);
            r[1] = static_cast<double>(s_score[i]);
            // warp_boxes: clip against the original size, one float32 round trip, then divide by the scale factor
            r[2] = static_cast<double>(static_cast<float>(fmin(fmax(x0, 0.0), static_cast<double>(org_w)))) / static_cast<double>(ratio_w);
            r[3] = static_cast<double>(static_cast<float>(fmin(fmax(y0, 0.0), static_cast<double>(org_h)))) / static_cast<double>(ratio_h);
            r[4] = static_cast<double>(static_cast<float>(fmin(fmax(x1, 0.0), static_cast<double>(org_w)))) / static_cast<double>(ratio_w);
            r[5] = static_cast<double>(static_cast<float>(fmin(fmax(y1, 0.0), static_cast<double>(org_h)))) / static_cast<double>(ratio_h);
        }
        ++picked;
        const double a1 = area_of(x0, y0, x1, y1);
        for (int j = i + 1 + threadIdx.x; j < m; j += blockDim.x) {
            if (s_dead[j]) continue;
            const double ov = area_of(fmax(s_box[j][0], x0), fmax(s_box[j][1], y0), fmin(s_box[j][2], x1), fmin(s_box[j][3], y1));
            const double a0 = area_of(s_box[j][0], s_box[j][1], s_box[j][2], s_box[j][3]);
            const double iou = ov / (a0 + a1 - ov + 1e-5);
            if (!(iou <= iou_thr)) s_dead[j] = 1;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) cls_n[n * kMaxClasses + c] = picked;
}

__global__ void k_pico_pack(int C, int out_cap, const double* __restrict__ cls_rows, const int* __restrict__ cls_n, double* __restrict__ out,
                            int* __restrict__ counts) {
    const int n = blockIdx.x;
    int base = 0;
    for (int c = 0; c < C; ++c) {
        const int k = cls_n[n * kMaxClasses + c];
        for (int i = threadIdx.x; i < k * 6; i += blockDim.x) {
            const int row = base + i / 6;
            if (row < out_cap) out[(static_cast<size_t>(n) * out_cap + row) * 6 + i % 6] = cls_rows[((static_cast<size_t>(n) * kMaxClasses + c) * kMaxKeep) * 6 + i];
        }
        base += k;
    }
    if (threadIdx.x == 0) counts[n] = min(base, out_cap);
}

}  // namespace

int picodet_decode(Engine* e, const float* const* scores, const float* const* dfl, int N, int C, int reg_max, const int* strides, int in_h,
                   int in_w, const float* org_hw_host, const float* scale_host, float score_thr, double iou_thr, int nms_top_k,
                   int keep_top_k, int out_cap, double* out, int32_t* counts) {
    if (N == 0) return 0;
    if (!scores || !dfl || !strides || !org_hw_host || !scale_host || !out || !counts || N < 0)
        return set_err(e, DV_ERR_ARG, "picodet_decode: bad arguments");
    if (C <= 0 || C > kMaxClasses || reg_max < 1 || reg_max > 15 || nms_top_k <= 0 || nms_top_k > kTopK || keep_top_k <= 0 ||
        keep_top_k > kMaxKeep || out_cap <= 0)
        return set_err(e, DV_ERR_UNSUPPORTED, "picodet_decode: classes <= %d, reg_max <= 15, nms_top_k <= %d, keep_top_k <= %d", kMaxClasses,
                       kTopK, kMaxKeep);
    LevelPtrs lp;
    int max_hw = 0;
    for (int l = 0; l < kLevels; ++l) {
        if (!scores[l] || !dfl[l] || strides[l] <= 0) return set_err(e, DV_ERR_ARG, "picodet_decode: null level %d", l);
        lp.score[l] = scores[l];
        lp.dfl[l] = dfl[l];
        lp.stride[l] = strides[l];
        // np.arange(input / stride): ceil of the float quotient
        const int fh = (in_h + strides[l] - 1) / strides[l];
        lp.fw[l] = (in_w + strides[l] - 1) / strides[l];
        lp.hw[l] = fh * lp.fw[l];
        if (lp.hw[l] > kMaxAnchors) return set_err(e, DV_ERR_UNSUPPORTED, "picodet_decode: %d anchors on level %d (> %d)", lp.hw[l], l, kMaxAnchors);
        if (lp.hw[l] > max_hw) max_hw = lp.hw[l];
    }
    auto it = e->aux.find("picodet");
    if (it == e->aux.end()) it = e->aux.emplace("picodet", std::unique_ptr<Model>(new PicoWs())).first;
    PicoWs* ws = static_cast<PicoWs*>(it->second.get());
    if (ws->N < N) {
        for (void* p : ws->mem) cudaFree(p);
        ws->mem.clear();
        auto alloc = [&](void** p, size_t bytes) -> int {
            cudaError_t st = cudaMalloc(p, bytes);
            if (st != cudaSuccess) return set_err(e, DV_ERR_CUDA, "picodet_decode: cudaMalloc(%zu): %s", bytes, cudaGetErrorString(st));
            ws->mem.push_back(*p);
            return 0;
        };
        const size_t n = static_cast<size_t>(N);
        DV_TRY(alloc(reinterpret_cast<void**>(&ws->cand_box), n * kLevels * kTopK * 32));
        DV_TRY(alloc(reinterpret_cast<void**>(&ws->cand_anchor), n * kLevels * kTopK * 4));
        DV_TRY(alloc(reinterpret_cast<void**>(&ws->cand_n), n * kLevels * 4));
        DV_TRY(alloc(reinterpret_cast<void**>(&ws->cls_rows), n * kMaxClasses * kMaxKeep * 48));
        DV_TRY(alloc(reinterpret_cast<void**>(&ws->cls_n), n * kMaxClasses * 4));
        DV_TRY(alloc(reinterpret_cast<void**>(&ws->meta), n * 16));
        ws->N = N;
    }
    std::vector<float> meta(static_cast<size_t>(N) * 4);
    for (int i = 0; i < N; ++i) {
        meta[i * 4 + 0] = org_hw_host[i * 2];
        meta[i * 4 + 1] = org_hw_host[i * 2 + 1];
        meta[i * 4 + 2] = scale_host[i * 2];
        meta[i * 4 + 3] = scale_host[i * 2 + 1];
    }
    cudaStream_t s = e->stream;
    DV_CUDA(e, cudaMemcpyAsync(ws->meta, meta.data(), meta.size() * 4, cudaMemcpyHostToDevice, s));
    DV_CUDA(e, cudaStreamSynchronize(s));  // `meta` is a stack-lifetime staging buffer
    static DeviceOnce attr_once;
    if (attr_once.need(e->device)) {
        DV_CUDA(e, cudaFuncSetAttribute(k_pico_select, cudaFuncAttributeMaxDynamicSharedMemorySize, kMaxAnchors * 8));
        attr_once.mark(e->device);
    }
    int n2 = 1;
    while (n2 < max_hw) n2 <<= 1;
    const int bins = reg_max + 1;
    double anchors = 0;
    for (int l = 0; l < kLevels; ++l) anchors += lp.hw[l];
    e->launch_begin("k_pico_select", "picodet_decode", 0.0, static_cast<double>(N) * anchors * C * 4.0);
    k_pico_select<<<dim3(kLevels, N), 256, static_cast<size_t>(n2) * 8, s>>>(lp, C, bins, score_thr, nms_top_k, ws->cand_box, ws->cand_anchor,
                                                                         ws->cand_n);
    e->launch_end();
    e->launch_begin("k_pico_nms", "picodet_decode", 0.0, static_cast<double>(N) * C * kNmsCand * 40.0);
    k_pico_nms<<<dim3(C, N), 256, 0, s>>>(lp, C, score_thr, iou_thr, keep_top_k, ws->cand_box, ws->cand_anchor, ws->cand_n,
                                          ws->meta, ws->cls_rows, ws->cls_n);
    e->launch_end();
    e->launch_begin("k_pico_pack", "picodet_decode", 0.0, static_cast<double>(N) * out_cap * 48.0);
    k_pico_pack<<<N, 128, 0, s>>>(C, out_cap, ws->cls_rows, ws->cls_n, out, counts);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

}  // namespace dv
