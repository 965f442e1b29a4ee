// ctc_greedy_decode: per time-step arg-max / max over the class axis, collapse repeats,
// drop blank, mean confidence of the kept steps.
// Follows the reference CTCLabelDecode.__call__ (ocr_rec_pp/rec_postprocess.py:175-191) and
// BaseRecLabelDecode.decode (:126-161):
//   idx = preds.argmax(axis=2) (first maximum wins), prob = preds.max(axis=2)
//   keep[t] = (t == 0 || idx[t] != idx[t-1]) && idx[t] != blank
//   conf    = np.mean(prob[keep]) in float32 (numpy pairwise summation order), 0 if nothing kept.
// One CTA per crop: warps stride over the T rows with 128-bit loads (one pass over the [T,C] slab,
// the only HBM traffic), then warp 0 compacts with ballots.
#include "engine.h"

namespace dv {

namespace {

constexpr int kMaxT = 1024;

struct ArgMax {
    float v;
    int i;
};
__device__ __forceinline__ ArgMax better(ArgMax a, ArgMax b) {
    // numpy argmax: first occurrence of the maximum
    if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
    return a;
}

// numpy's float32 pairwise summation (numpy/_core/src/umath/loops_utils.h.src, *_pairwise_sum)
__device__ float np_pairwise_sum(const float* a, int n) {
    if (n < 8) {
        float res = 0.f;
        for (int i = 0; i < n; ++i) res = __fadd_rn(res, a[i]);
        return res;
    } else if (n <= 128) {
        float r[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) r[j] = a[j];
        int i;
        for (i = 8; i < n - (n % 8); i += 8) {
#pragma unroll
            for (int j = 0; j < 8; ++j) r[j] = __fadd_rn(r[j], a[i + j]);
        }
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), __fadd_rn(r[2], r[3])),
                              __fadd_rn(__fadd_rn(r[4], r[5]), __fadd_rn(r[6], r[7])));
        for (; i < n; ++i) res = __fadd_rn(res, a[i]);
        return res;
    } else {
        int n2 = n / 2;
        n2 -= n2 % 8;
        return __fadd_rn(np_pairwise_sum(a, n2), np_pairwise_sum(a + n2, n - n2));
    }
}

// Second phase, one warp: keep[t] = id[t] != blank && (t == 0 || id[t] != id[t-1]); left-pack with ballots.
__device__ void collapse_write(const int* s_id, const float* s_p, float* s_kept, int T, int blank, int b, int lane,
                               int32_t* __restrict__ out_ids, int32_t* __restrict__ out_len,
                               float* __restrict__ out_conf) {
    int count = 0;
    for (int t0 = 0; t0 < T; t0 += 32) {
        const int t = t0 + lane;
        bool keep = false;
        int id = 0;
        if (t < T) {
            id = s_id[t];
            keep = (id != blank) && (t == 0 || id != s_id[t - 1]);
        }
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (keep) {
            const int pos = count + __popc(mask & ((1u << lane) - 1u));
            out_ids[static_cast<long long>(b) * T + pos] = id;
            s_kept[pos] = s_p[t];
        }
        count += __popc(mask);
    }
    for (int t = count + lane; t < T; t += 32) out_ids[static_cast<long long>(b) * T + t] = -1;
    __syncwarp();
    if (lane == 0) {
        out_len[b] = count;
        out_conf[b] = count > 0 ? __fdiv_rn(np_pairwise_sum(s_kept, count), static_cast<float>(count)) : 0.f;
    }
}

__global__ void __launch_bounds__(256)
k_ctc_greedy(const float* __restrict__ probs, int T, int C, int blank, int32_t* __restrict__ out_ids,
             int32_t* __restrict__ out_len, float* __restrict__ out_conf, int32_t* __restrict__ raw_ids,
             float* __restrict__ raw_max) {
    __shared__ int s_id[kMaxT];
    __shared__ float s_p[kMaxT];
    __shared__ float s_kept[kMaxT];
    const int b = blockIdx.x;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const float* base = probs + static_cast<long long>(b) * T * C;
    for (int t = warp; t < T; t += nwarps) {
        const float* row = base + static_cast<long long>(t) * C;
        ArgMax m{-INFINITY, 0x7fffffff};
        // scalar head up to 16-byte alignment, float4 body, scalar tail
        int head = static_cast<int>(((16u - (reinterpret_cast<uintptr_t>(row) & 15u)) & 15u) >> 2);
        if (head > C) head = C;
        if (lane < head) m = better(m, ArgMax{__ldg(row + lane), lane});
        const int nvec = (C - head) >> 2;
        const float4* rv = reinterpret_cast<const float4*>(row + head);
        for (int i = lane; i < nvec; i += 32) {
            const float4 f = __ldg(rv + i);
            const int c = head + i * 4;
            m = better(m, ArgMax{f.x, c});
            m = better(m, ArgMax{f.y, c + 1});
            m = better(m, ArgMax{f.z, c + 2});
            m = better(m, ArgMax{f.w, c + 3});
        }
        const int tail0 = head + nvec * 4;
        if (tail0 + lane < C) m = better(m, ArgMax{__ldg(row + tail0 + lane), tail0 + lane});
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            ArgMax other{__shfl_xor_sync(0xffffffffu, m.v, o), __shfl_xor_sync(0xffffffffu, m.i, o)};
            m = better(m, other);
        }
        if (lane == 0) {
            s_id[t] = m.i;
            s_p[t] = m.v;
            if (raw_ids) raw_ids[static_cast<long long>(b) * T + t] = m.i;
            if (raw_max) raw_max[static_cast<long long>(b) * T + t] = m.v;
        }
    }
    __syncthreads();
    if (warp != 0) return;
    collapse_write(s_id, s_p, s_kept, T, blank, b, lane, out_ids, out_len, out_conf);
}


// Collapse of precomputed per-step arg-max ids (the classifier's fused arg-max epilogue already reduced
// over the class axis).  Follows OCRRecognitionPostProcessor.__call__ (ocr_recognition/
// processor_ocr_recognition.py:152-162: `if p != last_p and p != 0`, last_p starting at 0), which is the CTC
// rule with blank = 0.  scores may be NULL (then out_conf = 0 for every row: the reference returns no
// confidence on this path).
__global__ void __launch_bounds__(32)
k_collapse_ids(const int32_t* __restrict__ ids, const float* __restrict__ scores, int T, int blank,
               int32_t* __restrict__ out_ids, int32_t* __restrict__ out_len, float* __restrict__ out_conf) {
    __shared__ int s_id[kMaxT];
    __shared__ float s_p[kMaxT];
    __shared__ float s_kept[kMaxT];
    const int b = blockIdx.x, lane = threadIdx.x;
    for (int t = lane; t < T; t += 32) {
        s_id[t] = ids[static_cast<long long>(b) * T + t];
        s_p[t] = scores ? scores[static_cast<long long>(b) * T + t] : 0.f;
    }
    __syncwarp();
    collapse_write(s_id, s_p, s_kept, T, blank, b, lane, out_ids, out_len, out_conf);
}

}  // namespace

int ctc_greedy(Engine* e, const float* probs, int B, int T, int C, int blank, int32_t* out_ids, int32_t* out_len,
               float* out_conf, int32_t* raw_ids, float* raw_max) {
    if (B == 0) return 0;
    if (!probs || !out_ids || !out_len || !out_conf || B < 0 || T <= 0 || C <= 0)
        return set_err(e, DV_ERR_ARG, "ctc_greedy: bad arguments");
    if (T > kMaxT) return set_err(e, DV_ERR_UNSUPPORTED, "ctc_greedy: T=%d > %d", T, kMaxT);
    // algorithmic bytes: the [B,T,C] fp32 slab read once + ids/len/conf written once
    e->launch_begin("k_ctc_greedy", "ctc", 0.0, 4.0 * B * T * C + 4.0 * B * T + 8.0 * B);
    k_ctc_greedy<<<B, 256, 0, e->stream>>>(probs, T, C, blank, out_ids, out_len, out_conf, raw_ids, raw_max);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

int ctc_collapse(Engine* e, const int32_t* ids, const float* scores, int B, int T, int blank, int32_t* out_ids,
                 int32_t* out_len, float* out_conf) {
    if (B == 0) return 0;
    if (!ids || !out_ids || !out_len || !out_conf || B < 0 || T <= 0) return set_err(e, DV_ERR_ARG, "ctc_collapse: bad arguments");
    if (T > kMaxT) return set_err(e, DV_ERR_UNSUPPORTED, "ctc_collapse: T=%d > %d", T, kMaxT);
    e->launch_begin("k_collapse_ids", "collapse", 0.0, (scores ? 8.0 : 4.0) * B * T + 4.0 * B * T + 8.0 * B);
    k_collapse_ids<<<B, 32, 0, e->stream>>>(ids, scores, T, blank, out_ids, out_len, out_conf);
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

}  // namespace dv
