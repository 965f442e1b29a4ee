// Lore logical-location processor: two 4-layer pre-norm transformer regressors over the selected table cells.
// Architecture restated from the reference LoreProcessModel.forward lore/lore_processor.py:465-514 (evaluation branch):
// Transformer :81-114, Encoder :39-61 (positional encoder and the final Norm are never applied), EncoderLayer :286-313,
// MultiHeadAttention :172-226 (8 heads, d_k = 32), Norm :117-131 (UNBIASED std, eps added to the std), FeedForward
// :229-242, Decoder :64-78 (ends in ReLU), Stacker :342-396.  The CPU mirror is oracle/lore_processor_ref.py.
//
// Numerics.  The outputs are rounded to integer row / column indices downstream (process_logic_output,
// lineless_table_process.py:658-663), so this model runs at ~fp32 accuracy on the tensor cores: every Linear is a
// split-fp16 GEMM on conv_igemm_tcgen05 (activations stored as [hi | lo] fp16 pairs, weights as [W_hi | W_lo | W_hi],
// three k-passes accumulated in fp32 TMEM -- csrc/igemm_host.cu plan_linear), LayerNorm / softmax / residual stream are
// fp32.  The cell count is data dependent: every GEMM reads its row count from device memory (m_dyn) and the attention
// kernel works on per-image segments given by device offsets, so the whole chain is enqueued without a host round trip.
#include <math.h>

#include "engine.h"

namespace dv {

namespace {

constexpr int kD = 256, kHeads = 8, kDk = 32, kFF = 2048;

// fp32 rows -> split fp16 [hi | lo]: out[r][c] = fp16(x), out[r][lo_off + c] = fp16(x - hi); columns C..Cpad zero
__global__ void __launch_bounds__(256)
k_split_rows(const float* __restrict__ in, int ld_in, int C, int Cpad, const int* __restrict__ m_dyn, int cap, __half* __restrict__ out,
             int ld_out, int lo_off) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const int rows = min(*m_dyn, cap);
    const long long r = idx / Cpad;
    if (r >= rows) return;
    const int c = static_cast<int>(idx % Cpad);
    const float v = c < C ? in[r * ld_in + c] : 0.f;
    const __half hi = __float2half_rn(v);
    out[r * ld_out + c] = hi;
    out[r * ld_out + lo_off + c] = __float2half_rn(v - __half2float(hi));
}

// Norm (lore_processor.py:117-131): alpha * (x - mean) / (std_unbiased + eps) + bias, one warp per 256-wide row
__global__ void __launch_bounds__(256)
k_norm_split(const float* __restrict__ x, const float* __restrict__ alpha, const float* __restrict__ bias, const int* __restrict__ m_dyn,
             int cap, __half* __restrict__ out) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (r >= min(*m_dyn, cap)) return;
    float v[8];
    const float4* p = reinterpret_cast<const float4*>(x + static_cast<long long>(r) * kD + lane * 8);
    const float4 a = p[0], b = p[1];
    v[0] = a.x, v[1] = a.y, v[2] = a.z, v[3] = a.w, v[4] = b.x, v[5] = b.y, v[6] = b.z, v[7] = b.w;
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    const float mean = s / kD;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        v[i] -= mean;
        q += v[i] * v[i];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
    const float inv = 1.f / (sqrtf(q / (kD - 1)) + 1e-6f);
    __half hi[8], lo[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = lane * 8 + i;
        const float y = alpha[c] * v[i] * inv + bias[c];
        hi[i] = __float2half_rn(y);
        lo[i] = __float2half_rn(y - __half2float(hi[i]));
    }
    __half* o = out + static_cast<long long>(r) * (2 * kD) + lane * 8;
    *reinterpret_cast<uint4*>(o) = *reinterpret_cast<const uint4*>(hi);
    *reinterpret_cast<uint4*>(o + kD) = *reinterpret_cast<const uint4*>(lo);
}

// Self-attention inside each image's segment of cells.  qkv fp32 [rows, 768] (q pre-scaled by 1/sqrt(d_k)).
// One warp per (query row, head): lane l scores keys l, l+32, ... with an online softmax, then the 32 partial
// (max, sum, acc[32]) states are merged.  Output: split fp16 [rows, 512].
__global__ void __launch_bounds__(256)
k_attn_seg(const float* __restrict__ qkv, const int32_t* __restrict__ offsets, int n_img, const int* __restrict__ m_dyn, int cap,
           __half* __restrict__ out) {
    const int r = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31, h = blockIdx.y;
    if (r >= min(*m_dyn, cap)) return;
    int img = 0;
    while (img + 1 < n_img && offsets[img + 1] <= r) ++img;
    const int s0 = offsets[img], s1 = min(offsets[img + 1], cap);
    float q[kDk];
    {
        const float4* qp = reinterpret_cast<const float4*>(qkv + static_cast<long long>(r) * 768 + h * kDk);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 t = __ldg(qp + i);
            q[4 * i] = t.x, q[4 * i + 1] = t.y, q[4 * i + 2] = t.z, q[4 * i + 3] = t.w;
        }
    }
    float mx = -INFINITY, sum = 0.f, acc[kDk];
#pragma unroll
    for (int i = 0; i < kDk; ++i) acc[i] = 0.f;
    for (int j = s0 + lane; j < s1; j += 32) {
        const float4* kp = reinterpret_cast<const float4*>(qkv + static_cast<long long>(j) * 768 + 256 + h * kDk);
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 t = __ldg(kp + i);
            s = fmaf(q[4 * i], t.x, s);
            s = fmaf(q[4 * i + 1], t.y, s);
            s = fmaf(q[4 * i + 2], t.z, s);
            s = fmaf(q[4 * i + 3], t.w, s);
        }
        const float nm = fmaxf(mx, s);
        const float corr = expf(mx - nm), pj = expf(s - nm);
        sum = sum * corr + pj;
        const float4* vp = reinterpret_cast<const float4*>(qkv + static_cast<long long>(j) * 768 + 512 + h * kDk);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float4 t = __ldg(vp + i);
            acc[4 * i] = fmaf(pj, t.x, acc[4 * i] * corr);
            acc[4 * i + 1] = fmaf(pj, t.y, acc[4 * i + 1] * corr);
            acc[4 * i + 2] = fmaf(pj, t.z, acc[4 * i + 2] * corr);
            acc[4 * i + 3] = fmaf(pj, t.w, acc[4 * i + 3] * corr);
        }
        mx = nm;
    }
    float gm = mx;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) gm = fmaxf(gm, __shfl_xor_sync(0xffffffffu, gm, o));
    const float sc = mx == -INFINITY ? 0.f : expf(mx - gm);
    sum *= sc;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    float mine = 0.f;
#pragma unroll
    for (int d = 0; d < kDk; ++d) {
        float v = acc[d] * sc;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == d) mine = v;
    }
    const float y = mine / sum;
    const __half hi = __float2half_rn(y);
    __half* o = out + static_cast<long long>(r) * (2 * kD) + h * kDk + lane;
    o[0] = hi;
    o[kD] = __float2half_rn(y - __half2float(hi));
}

struct PStep {
    enum Kind { GEMM, SPLIT, NORM, ATTN } kind;
    ConvPlan plan;
    const float *fin = nullptr, *alpha = nullptr, *bias = nullptr;
    __half* hout = nullptr;
    int ld_in = 0, C = 0, Cpad = 0, ld_out = 0, lo_off = 0;
    std::string name;
};

struct LoreProc : Model {
    Engine* e = nullptr;
    int cap = 0, n_img_cap = 0;
    int n_axis = 4, n_stack = 4;
    std::vector<void*> mem;
    std::vector<PStep> steps;
    float *feat = nullptr, *xs = nullptr, *qkv = nullptr, *logic = nullptr, *stacked = nullptr;
    __half *in_split = nullptr, *a512 = nullptr, *hsplit = nullptr, *dec = nullptr, *lsplit = nullptr, *e1 = nullptr, *cat = nullptr;
    int32_t *rows_dev = nullptr, *offsets = nullptr;
    double flops = 0;
    ~LoreProc() override {
        for (void* p : mem) cudaFree(p);
    }
    template <typename T>
    int alloc(T** p, size_t elems) {
        void* q = nullptr;
        cudaError_t st = cudaMalloc(&q, elems * sizeof(T) + 16);
        if (st != cudaSuccess) return set_err(e, DV_ERR_CUDA, "lore_proc: cudaMalloc failed: %s", cudaGetErrorString(st));
        cudaMemsetAsync(q, 0, elems * sizeof(T), e->stream);
        mem.push_back(q);
        *p = reinterpret_cast<T*>(q);
        return 0;
    }
};

int get_split_linear(Engine* e, const std::string& name, int K, int N, ConvSpec* cs) {
    const BlobTensor* w = e->find(name + ".w");
    const BlobTensor* b = e->find(name + ".b");
    if (!w || !b || w->dtype != 1 || b->dtype != 0 || w->ndim != 2) return set_err(e, DV_ERR_WEIGHTS, "missing weights for '%s'", name.c_str());
    const int Kp = K >= 64 ? (K + 63) / 64 * 64 : (K + 15) / 16 * 16;
    if (static_cast<int>(w->dims[0]) != N || static_cast<int>(w->dims[1]) != 3 * Kp)
        return set_err(e, DV_ERR_WEIGHTS, "'%s': weight [%u,%u] != [%d,3*%d]", name.c_str(), w->dims[0], w->dims[1], N, Kp);
    cs->KH = cs->KW = 1;
    cs->Cin = Kp;
    cs->Cin_pad = 3 * Kp;
    cs->Cout = N;
    cs->BK = Kp >= 64 ? 64 : 16;
    cs->w = reinterpret_cast<const __half*>(w->dptr);
    cs->bias = reinterpret_cast<const float*>(b->dptr);
    cs->flat = true;
    cs->split = true;
    return 0;
}

// A: split fp16 [cap, lda] with hi at column 0 and lo at column Kp
int add_gemm(LoreProc* m, const std::string& name, const __half* A, int lda, int K, int N, EpiSpec es) {
    ConvSpec cs;
    DV_TRY(get_split_linear(m->e, name, K, N, &cs));
    es.m_dyn = m->rows_dev;
    PStep st;
    st.kind = PStep::GEMM;
    st.name = name;
    // the planner's split layout puts lo at column K of a 2K-wide row; wider rows (lda) keep lo at lda/2
    if (lda != 2 * cs.Cin) return set_err(m->e, DV_ERR_ARG, "%s: split operand must be [hi | lo] of width 2*%d (got %d)", name.c_str(), cs.Cin, lda);
    DV_TRY(plan_linear(m->e, A, m->cap, cs.Cin, cs, es, &st.plan, st.name.c_str(), lda));
    m->mem.push_back(m->e->owned.back());
    m->e->owned.pop_back();
    m->flops += st.plan.flops;
    m->steps.push_back(st);
    return 0;
}

EpiSpec to_stream(float* xs, bool residual) {
    EpiSpec es;
    es.out = xs;
    es.out_ld = kD;
    es.out_f32 = 1;
    if (residual) {
        es.res = xs;
        es.res_mode = RES_SAME;
        es.res_ld = kD;
        es.res_f32 = 1;
    }
    return es;
}

EpiSpec to_split(__half* out, int K, int act) {
    EpiSpec es;
    es.out = out;
    es.out_ld = 2 * K;
    es.act = act;
    es.split_off = K;
    return es;
}

int add_split(LoreProc* m, const float* in, int ld_in, int C, int Cpad, __half* out, int ld_out, int lo_off, const char* name) {
    PStep st;
    st.kind = PStep::SPLIT;
    st.fin = in;
    st.ld_in = ld_in;
    st.C = C;
    st.Cpad = Cpad;
    st.hout = out;
    st.ld_out = ld_out;
    st.lo_off = lo_off;
    st.name = name;
    m->steps.push_back(st);
    return 0;
}

int add_transformer(LoreProc* m, const std::string& p, const __half* in_split, int K_in, int n_layers, float* out4) {
    Engine* e = m->e;
    DV_TRY(add_gemm(m, p + ".in", in_split, 2 * K_in, K_in, kD, to_stream(m->xs, false)));
    for (int L = 0; L < n_layers; ++L) {
        const std::string lp = p + "." + std::to_string(L);
        for (int half = 0; half < 2; ++half) {
            const std::string nn = lp + (half == 0 ? ".norm_1" : ".norm_2");
            const BlobTensor* a = e->find(nn + ".a");
            const BlobTensor* b = e->find(nn + ".b");
            if (!a || !b || a->nbytes != kD * 4 || b->nbytes != kD * 4) return set_err(e, DV_ERR_WEIGHTS, "missing '%s'", nn.c_str());
            PStep st;
            st.kind = PStep::NORM;
            st.alpha = reinterpret_cast<const float*>(a->dptr);
            st.bias = reinterpret_cast<const float*>(b->dptr);
            st.name = nn;
            m->steps.push_back(st);
            if (half == 0) {
                EpiSpec es;
                es.out = m->qkv;
                es.out_ld = 3 * kD;
                es.out_f32 = 1;
                DV_TRY(add_gemm(m, lp + ".qkv", m->a512, 2 * kD, kD, 3 * kD, es));
                PStep at;
                at.kind = PStep::ATTN;
                at.name = lp + ".attn";
                m->steps.push_back(at);
                DV_TRY(add_gemm(m, lp + ".out", m->a512, 2 * kD, kD, kD, to_stream(m->xs, true)));
            } else {
                DV_TRY(add_gemm(m, lp + ".ff1", m->a512, 2 * kD, kD, kFF, to_split(m->hsplit, kFF, ACT_RELU)));
                DV_TRY(add_gemm(m, lp + ".ff2", m->hsplit, 2 * kFF, kFF, kD, to_stream(m->xs, true)));
            }
        }
    }
    DV_TRY(add_split(m, m->xs, kD, kD, kD, m->a512, 2 * kD, kD, (p + ".dec.split").c_str()));
    DV_TRY(add_gemm(m, p + ".dec0", m->a512, 2 * kD, kD, kD, to_split(m->dec, kD, ACT_RELU)));
    EpiSpec es;
    es.out = out4;
    es.out_ld = 4;
    es.out_f32 = 1;
    es.act = ACT_RELU;
    DV_TRY(add_gemm(m, p + ".dec2", m->dec, 2 * kD, kD, 4, es));
    return 0;
}

int build(Engine* e, LoreProc* m, int cap, int n_img_cap) {
    for (void* p : m->mem) cudaFree(p);
    m->mem.clear();
    m->steps.clear();
    m->flops = 0;
    m->e = e;
    m->cap = cap;
    m->n_img_cap = n_img_cap;
    const BlobTensor* meta = e->find("meta");
    if (meta && meta->dtype == 2 && meta->nbytes >= 16) {
        int32_t h[4];
        DV_CUDA(e, cudaMemcpy(h, meta->dptr, 16, cudaMemcpyDeviceToHost));
        m->n_axis = h[0];
        m->n_stack = h[1];
    }
    const size_t c = static_cast<size_t>(cap);
    DV_TRY(m->alloc(&m->rows_dev, 4));
    DV_TRY(m->alloc(&m->offsets, static_cast<size_t>(n_img_cap) + 1));
    DV_TRY(m->alloc(&m->xs, c * kD));
    DV_TRY(m->alloc(&m->qkv, c * 3 * kD));
    DV_TRY(m->alloc(&m->logic, c * 4));
    DV_TRY(m->alloc(&m->stacked, c * 4));
    DV_TRY(m->alloc(&m->in_split, c * 2 * kD));
    DV_TRY(m->alloc(&m->a512, c * 2 * kD));
    DV_TRY(m->alloc(&m->hsplit, c * 2 * kFF));
    DV_TRY(m->alloc(&m->dec, c * 2 * kD));
    DV_TRY(m->alloc(&m->lsplit, c * 32));
    DV_TRY(m->alloc(&m->e1, c * 2 * kD));
    DV_TRY(m->alloc(&m->cat, c * 4 * kD));
    // base regressor: feat -> logic
    DV_TRY(add_split(m, nullptr /* feat, set per call */, kD, kD, kD, m->in_split, 2 * kD, kD, "feat.split"));
    DV_TRY(add_transformer(m, "axis", m->in_split, kD, m->n_axis, m->logic));
    // stacker: cat(logi_encoder(logic), feat) -> stacked   (cat as split: hi = [emb | feat], lo = [emb_lo | feat_lo])
    DV_TRY(add_split(m, m->logic, 4, 4, 16, m->lsplit, 32, 16, "logic.split"));
    DV_TRY(add_gemm(m, "stack.enc0", m->lsplit, 32, 4, kD, to_split(m->e1, kD, ACT_RELU)));
    {
        EpiSpec es = to_split(m->cat, 2 * kD, ACT_RELU);  // hi -> cat[:, 0:256], lo -> cat[:, 512:768]
        DV_TRY(add_gemm(m, "stack.enc2", m->e1, 2 * kD, kD, kD, es));
    }
    DV_TRY(add_split(m, nullptr /* feat */, kD, kD, kD, m->cat + kD, 4 * kD, 2 * kD, "feat.cat"));
    DV_TRY(add_transformer(m, "stack", m->cat, 2 * kD, m->n_stack, m->stacked));
    return 0;
}

}  // namespace

int lore_proc_create(Engine* e) {
    LoreProc* m = new LoreProc();
    m->e = e;
    e->model.reset(m);
    return 0;
}

// 2-D position embeddings of the wiz_2dpe configurations (ptn, wireless; LoreProcessModel.forward, lore_processor.py:486-490):
// feat[row] = feat[row] + x_pe[d0] + y_pe[d1] + x_pe[d2] + y_pe[d5] (fp32, the reference's left-to-right order), row =
// offsets[n] + j for cell j of image n, d = dets_feat[n][j][0..7] (the integer position features of the decode), clamped to the
// table of 256 positions.  In place; one thread per (row, 4 channels).
__global__ void __launch_bounds__(256)
k_add_pos_emb(float* __restrict__ feat, const int32_t* __restrict__ dets, const int32_t* __restrict__ counts, const int32_t* __restrict__ offsets,
              int K, int cap, const float* __restrict__ xpe, const float* __restrict__ ype) {
    const int n = blockIdx.y, j = blockIdx.x * 4 + (threadIdx.x >> 6), c4 = threadIdx.x & 63;
    if (j >= counts[n]) return;
    const int row = offsets[n] + j;
    if (row >= cap) return;
    const int32_t* d = dets + (static_cast<long long>(n) * K + j) * 8;
    auto pos = [](int v) { return min(max(v, 0), 255); };
    const float4 a = *reinterpret_cast<const float4*>(xpe + pos(d[0]) * kD + c4 * 4), b = *reinterpret_cast<const float4*>(ype + pos(d[1]) * kD + c4 * 4);
    const float4 c = *reinterpret_cast<const float4*>(xpe + pos(d[2]) * kD + c4 * 4), e = *reinterpret_cast<const float4*>(ype + pos(d[5]) * kD + c4 * 4);
    float4* fp = reinterpret_cast<float4*>(feat + static_cast<long long>(row) * kD + c4 * 4);
    float4 f = *fp;
    f.x = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(f.x, a.x), b.x), c.x), e.x);
    f.y = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(f.y, a.y), b.y), c.y), e.y);
    f.z = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(f.z, a.z), b.z), c.z), e.z);
    f.w = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(f.w, a.w), b.w), c.w), e.w);
    *fp = f;
}

int lore_add_position_embeddings(Engine* e, float* feat, int cap_rows, const int32_t* dets_feat, const int32_t* counts, const int32_t* offsets,
                                 int n_img, int K) {
    if (!dynamic_cast<LoreProc*>(e->model.get())) return set_err(e, DV_ERR_STATE, "handle was not created as a lore_processor model");
    if (!feat || !dets_feat || !counts || !offsets || cap_rows <= 0 || n_img <= 0 || K <= 0)
        return set_err(e, DV_ERR_ARG, "lore_add_position_embeddings: bad arguments");
    const BlobTensor* xp = e->find("x_pos");
    const BlobTensor* yp = e->find("y_pos");
    if (!xp || !yp || xp->dtype != 0 || yp->dtype != 0 || xp->nbytes < 256 * kD * 4 || yp->nbytes < 256 * kD * 4)
        return set_err(e, DV_ERR_WEIGHTS, "lore_processor: missing x_pos / y_pos embedding tables [256,256]");
    e->launch_begin("k_add_pos_emb", "pos_emb", 0.0, static_cast<double>(cap_rows) * kD * 4.0 * 6.0);
    k_add_pos_emb<<<dim3((K + 3) / 4, n_img), 256, 0, e->stream>>>(feat, dets_feat, counts, offsets, K, cap_rows, reinterpret_cast<const float*>(xp->dptr),
                                                                  reinterpret_cast<const float*>(yp->dptr));
    e->launch_end();
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

// feat fp32 [cap_rows, 256] (device), rows_dev: device int = number of valid rows, offsets: device int32 [n_img + 1]
// (per-image segments of the rows), logic_out / stacked_out fp32 [cap_rows, 4] (device).
int lore_process_forward(Engine* e, const float* feat, int cap_rows, const int32_t* rows_dev, const int32_t* offsets, int n_img,
                         float* logic_out, float* stacked_out) {
    LoreProc* m = dynamic_cast<LoreProc*>(e->model.get());
    if (!m) return set_err(e, DV_ERR_STATE, "handle was not created as a lore_processor model");
    if (!feat || !rows_dev || !offsets || !stacked_out || cap_rows <= 0 || n_img <= 0)
        return set_err(e, DV_ERR_ARG, "lore_process_forward: bad arguments");
    if (m->cap != cap_rows || m->n_img_cap < n_img) DV_TRY(build(e, m, cap_rows, n_img));
    cudaStream_t s = e->stream;
    DV_CUDA(e, cudaMemcpyAsync(m->rows_dev, rows_dev, 4, cudaMemcpyDeviceToDevice, s));
    DV_CUDA(e, cudaMemcpyAsync(m->offsets, offsets, (n_img + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice, s));
    const int cap = m->cap;
    for (PStep& st : m->steps) {
        switch (st.kind) {
            case PStep::GEMM: DV_TRY(launch_conv(e, st.plan)); break;
            case PStep::SPLIT: {
                const float* in = st.fin ? st.fin : feat;
                const long long total = static_cast<long long>(cap) * st.Cpad;
                e->launch_begin("k_split_rows", st.name, 0.0, total * 8.0);
                k_split_rows<<<static_cast<int>((total + 255) / 256), 256, 0, s>>>(in, st.ld_in, st.C, st.Cpad, m->rows_dev, cap, st.hout,
                                                                                  st.ld_out, st.lo_off);
                e->launch_end();
                break;
            }
            case PStep::NORM:
                e->launch_begin("k_norm_split", st.name, 0.0, static_cast<double>(cap) * kD * 8.0);
                k_norm_split<<<(cap + 7) / 8, 256, 0, s>>>(m->xs, st.alpha, st.bias, m->rows_dev, cap, m->a512);
                e->launch_end();
                break;
            case PStep::ATTN:
                e->launch_begin("k_attn_seg", st.name, 4.0 * cap * cap * kD, static_cast<double>(cap) * 3 * kD * 4.0);
                k_attn_seg<<<dim3((cap + 7) / 8, kHeads), 256, 0, s>>>(m->qkv, m->offsets, n_img, m->rows_dev, cap, m->a512);
                e->launch_end();
                break;
        }
    }
    DV_CUDA(e, cudaGetLastError());
    const size_t bytes = static_cast<size_t>(cap) * 4 * sizeof(float);
    if (logic_out) DV_CUDA(e, cudaMemcpyAsync(logic_out, m->logic, bytes, cudaMemcpyDeviceToDevice, s));
    DV_CUDA(e, cudaMemcpyAsync(stacked_out, m->stacked, bytes, cudaMemcpyDeviceToDevice, s));
    return 0;
}

}  // namespace dv
