// CRNN text-line recogniser (north_star "SVTR/CRNN text-line recognition") as a static plan of launches.
// Architecture restated from the reference module crnn/modeling_crnn.py:36-113 (CRNN.forward): RGB -> gray, seven 3x3 convs
// + BatchNorm + ReLU with max-pools (2,2) (2,2) (2,1) (2,1), a (2,1) conv that folds the last two rows, two bidirectional
// LSTMs (hidden 256) each followed by a Linear, and the 512 -> 7644 classifier; output [b, w / 4, 7644] which
// OCRRecognitionPostProcessor (ocr_recognition/processor_ocr_recognition.py:147-165) soft-maxes, arg-maxes and collapses.
//
// Mapping to the engine:
//   * convs        conv_igemm_tcgen05 (implicit GEMM over NHWC fp16; conv0's single gray channel sits in an 8-channel pixel
//                  whose other channels are zero; conv4's (2,1) kernel is the A_PATCH mode with KH = 2, KW = 1, no padding)
//   * LSTM         input projections of both directions and all T steps as ONE flat GEMM (fp32 out, biases b_ih + b_hh
//                  folded); the recurrence as one small flat GEMM per step and direction -- A = h(t-1) [M, 256] fp16,
//                  W_hh [1024, 256], the projection row of step t added as an fp32 residual in the epilogue -- followed by
//                  k_lstm_cell (gates -> c (fp32), h (fp16: next step's operand and the layer's output sequence)).  The
//                  step GEMMs ping-pong between two h buffers, so each direction needs two plans whose residual pointer is
//                  patched per step.  Sequential by nature: 2 launches x T steps x 2 layers, ~10 us each.
//   * classifier   the arg-max epilogue of conv_igemm_tcgen05 (ids + max logit; the logits only when asked for).
// Crops (or chunks) are processed in passes of at most `pass_n` images so that the fp32 projection buffer stays bounded.
#include <stdlib.h>

#include "engine.h"

namespace dv {

int op_maxpool2x2(Engine* e, const Tensor& in, const Tensor& out);

namespace {

constexpr int kHid = 256, kGates = 4 * kHid;

static inline int grid_for(long long n, int block) { return static_cast<int>((n + block - 1) / block); }

// fp32 NCHW [M,3,H,W] (already / 255) -> fp16 NHWC [M,H,W,8]: channel 0 = 0.2989 R + 0.5870 G + 0.1140 B evaluated like torch
// (three fp32 products summed left to right, no FMA contraction), channels 1..7 = 0
__global__ void __launch_bounds__(256)
k_crnn_gray(const float* __restrict__ in, long long npix, long long plane, __half* __restrict__ out) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= npix) return;
    const long long n = idx / plane, r = idx - n * plane;
    const float* ip = in + n * 3 * plane + r;
    const float g = __fadd_rn(__fadd_rn(__fmul_rn(ip[0], 0.2989f), __fmul_rn(ip[plane], 0.5870f)), __fmul_rn(ip[2 * plane], 0.1140f));
    uint4 u = make_uint4(0u, 0u, 0u, 0u);
    u.x = static_cast<uint32_t>(__half_as_ushort(__float2half_rn(g)));
    *reinterpret_cast<uint4*>(out + idx * 8) = u;
}

// MaxPool2d((2,1)): [M,H,W,C] -> [M,H/2,W,C], 8 channels per thread
__global__ void __launch_bounds__(256)
k_maxpool_h2(const __half* __restrict__ in, long long total8, int W, int C, __half* __restrict__ out) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= total8) return;
    const int cv = C >> 3;
    const long long rowlen = static_cast<long long>(W) * cv;  // uint4 per image row
    const long long orow = idx / rowlen, within = idx - orow * rowlen;
    const uint4* ip = reinterpret_cast<const uint4*>(in) + (2 * orow) * rowlen + within;
    const uint4 a = __ldg(ip), b = __ldg(ip + rowlen);
    const __half2 *ha = reinterpret_cast<const __half2*>(&a), *hb = reinterpret_cast<const __half2*>(&b);
    uint4 o;
    __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) ho[i] = __hmax2(ha[i], hb[i]);
    reinterpret_cast<uint4*>(out)[idx] = o;
}

// One LSTM step of both directions (torch gate order i, f, g, o): gates fp32 [2][M,1024] (W_hh h + W_ih x + b), c fp32 [2][M,256]
// in place, h fp16 [2][M,256] for the next step, y fp16 [M*T, 512]: forward half at step t, backward half at step T-1-t.
__global__ void __launch_bounds__(256)
k_lstm_cell(const float* __restrict__ gates, float* __restrict__ c, __half* __restrict__ h_next, __half* __restrict__ y, int M, int T, int t) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    const long long per_dir = static_cast<long long>(M) * kHid;
    if (idx >= 2 * per_dir) return;
    const int d = idx >= per_dir ? 1 : 0;
    const long long r = idx - d * per_dir;
    const int m = static_cast<int>(r / kHid), j = static_cast<int>(r - static_cast<long long>(m) * kHid);
    const float* g = gates + (static_cast<long long>(d) * M + m) * kGates;
    const float gi = 1.f / (1.f + expf(-g[j])), gf = 1.f / (1.f + expf(-g[kHid + j]));
    const float gg = tanhf(g[2 * kHid + j]), go = 1.f / (1.f + expf(-g[3 * kHid + j]));
    const float cn = fmaf(gf, c[idx], gi * gg);
    c[idx] = cn;
    const __half hv = __float2half_rn(go * tanhf(cn));
    h_next[idx] = hv;
    const int ts = d ? T - 1 - t : t;
    y[(static_cast<long long>(m) * T + ts) * (2 * kHid) + d * kHid + j] = hv;
}

struct Step {
    enum Kind { GRAY, CONV, POOL22, POOLH2, LSTM, GEMM } kind;
    ConvPlan plan;
    Tensor a, b;
    int layer = 0;
    std::string name;
};

struct LstmLayer {
    ConvPlan step[2][2];   // [direction][which h buffer is the operand]
    const float* xproj = nullptr;
    float *gates = nullptr, *c = nullptr;
    __half* h[2] = {nullptr, nullptr};  // [buffer] -> [2 dirs][M,256]
    __half* y = nullptr;
};

struct Pass {
    Engine* e = nullptr;
    int M = 0, W = 0, T = 0;
    std::vector<void*> mem;
    std::vector<Step> steps;
    LstmLayer lstm[2];
    int cls_step = -1;
    float* stage_in = nullptr;
    __half* gray = nullptr;
    double flops = 0;
    ~Pass() {
        for (void* p : mem) cudaFree(p);
    }
    int alloc(void** p, size_t bytes, bool zero = false) {
        cudaError_t st = cudaMalloc(p, bytes ? bytes : 16);
        if (st != cudaSuccess) return set_err(e, DV_ERR_CUDA, "crnn: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(st));
        mem.push_back(*p);
        if (zero) cudaMemset(*p, 0, bytes);
        return 0;
    }
    int tensor(Tensor* t, int n, int h, int w, int c) {
        t->N = n, t->H = h, t->W = w, t->C = c, t->ld = 0, t->lo = 0;
        void* p = nullptr;
        DV_TRY(alloc(&p, t->elems() * sizeof(__half)));
        t->p = reinterpret_cast<__half*>(p);
        return 0;
    }
    void adopt() {  // plan_* allocate their delta tables through the engine: move them into this pass's pool
        mem.push_back(e->owned.back());
        e->owned.pop_back();
    }
};

struct CrnnModel : Model {
    int labels = 0;
    int pass_n = 768;
    std::map<long long, std::unique_ptr<Pass>> passes;  // key = M * 4096 + W
    double last_flops = 0;
};

int get_weights(Engine* e, const std::string& name, int rows, ConvSpec* cs, bool need_bias = true) {
    const BlobTensor* w = e->find(name + ".w");
    const BlobTensor* b = e->find(name + ".b");
    if (!w || w->dtype != 1 || w->ndim != 2 || static_cast<int>(w->dims[0]) != rows || (need_bias && (!b || b->dtype != 0)))
        return set_err(e, DV_ERR_WEIGHTS, "crnn: missing / bad weights for '%s'", name.c_str());
    const int taps = cs->KH * cs->KW;
    if (w->dims[1] % taps) return set_err(e, DV_ERR_WEIGHTS, "crnn: '%s' weight width %u not a multiple of %d taps", name.c_str(), w->dims[1], taps);
    cs->Cout = rows;
    cs->Cin_pad = static_cast<int>(w->dims[1]) / taps;
    cs->BK = (cs->Cin_pad % 64 == 0) ? 64 : (cs->Cin_pad % 32 == 0) ? 32 : 16;
    cs->w = reinterpret_cast<const __half*>(w->dptr);
    cs->bias = b ? reinterpret_cast<const float*>(b->dptr) : nullptr;
    if (b && b->dims[0] < static_cast<uint32_t>((rows + 255) / 256 * 256)) return set_err(e, DV_ERR_WEIGHTS, "crnn: '%s' bias not padded to 256", name.c_str());
    return 0;
}

int add_conv(Engine* e, Pass* ps, const std::string& name, const Tensor& in, int cout, int kh, int kw, int pad, Tensor* out, int Ho, int Wo) {
    DV_TRY(ps->tensor(out, in.N, Ho, Wo, cout));
    ConvSpec cs;
    cs.KH = kh, cs.KW = kw, cs.stride = 1, cs.pad = pad, cs.Cin = in.C;
    DV_TRY(get_weights(e, name, cout, &cs));
    EpiSpec es;
    es.out = out->p, es.out_ld = cout, es.act = ACT_RELU;
    Step st;
    st.kind = Step::CONV;
    st.name = name;
    DV_TRY(plan_conv(e, in, cs, es, Ho, Wo, &st.plan, name.c_str()));
    ps->adopt();
    ps->flops += st.plan.flops;
    ps->steps.push_back(st);
    return 0;
}

int add_pool(Pass* ps, Step::Kind kind, const Tensor& in, Tensor* out) {
    DV_TRY(ps->tensor(out, in.N, in.H / 2, kind == Step::POOL22 ? in.W / 2 : in.W, in.C));
    Step st;
    st.kind = kind;
    st.a = in, st.b = *out;
    ps->steps.push_back(st);
    return 0;
}

int add_linear(Engine* e, Pass* ps, const std::string& name, const __half* A, long long rows, int K, int N, const EpiSpec& es, bool bias = true) {
    ConvSpec cs;
    cs.KH = cs.KW = 1;
    cs.Cin = K;
    cs.flat = true;
    DV_TRY(get_weights(e, name, N, &cs, bias));
    if (cs.Cin_pad != K) return set_err(e, DV_ERR_WEIGHTS, "crnn: '%s' K %d != %d", name.c_str(), cs.Cin_pad, K);
    Step st;
    st.kind = Step::GEMM;
    st.name = name;
    DV_TRY(plan_linear(e, A, static_cast<int>(rows), K, cs, es, &st.plan, name.c_str()));
    ps->adopt();
    ps->flops += st.plan.flops;
    ps->steps.push_back(st);
    return 0;
}

// BidirectionalLSTM (modeling_crnn.py:21-33): x [M*T, In] -> Linear(concat(h_fwd, h_bwd)) [M*T, Nout]
int add_bilstm(Engine* e, Pass* ps, int layer, const __half* x, int In, int Nout, __half** out) {
    const std::string p = "rnn." + std::to_string(layer);
    const long long rows = static_cast<long long>(ps->M) * ps->T;
    LstmLayer& L = ps->lstm[layer];
    void* q = nullptr;
    DV_TRY(ps->alloc(&q, rows * 2 * kGates * sizeof(float)));
    float* xproj = reinterpret_cast<float*>(q);
    L.xproj = xproj;
    {
        EpiSpec es;
        es.out = xproj, es.out_ld = 2 * kGates, es.out_f32 = 1;
        DV_TRY(add_linear(e, ps, p + ".ih", x, rows, In, 2 * kGates, es));
    }
    DV_TRY(ps->alloc(&q, static_cast<size_t>(2) * ps->M * kGates * sizeof(float)));
    L.gates = reinterpret_cast<float*>(q);
    DV_TRY(ps->alloc(&q, static_cast<size_t>(2) * ps->M * kHid * sizeof(float)));
    L.c = reinterpret_cast<float*>(q);
    for (int i = 0; i < 2; ++i) {
        DV_TRY(ps->alloc(&q, static_cast<size_t>(2) * ps->M * kHid * sizeof(__half)));
        L.h[i] = reinterpret_cast<__half*>(q);
    }
    DV_TRY(ps->alloc(&q, rows * 2 * kHid * sizeof(__half)));
    L.y = reinterpret_cast<__half*>(q);
    for (int d = 0; d < 2; ++d)
        for (int buf = 0; buf < 2; ++buf) {
            ConvSpec cs;
            cs.KH = cs.KW = 1;
            cs.Cin = kHid;
            cs.flat = true;
            const std::string wn = p + (d ? ".hh_reverse" : ".hh");
            DV_TRY(get_weights(e, wn, kGates, &cs, /*need_bias=*/false));
            cs.bias = nullptr;
            EpiSpec es;
            es.out = L.gates + static_cast<size_t>(d) * ps->M * kGates;
            es.out_ld = kGates;
            es.out_f32 = 1;
            es.res = xproj + d * kGates;  // patched per step: + ts * 2 * kGates
            es.res_mode = RES_SAME;
            es.res_ld = ps->T * 2 * kGates;
            es.res_f32 = 1;
            DV_TRY(plan_linear(e, L.h[buf] + static_cast<size_t>(d) * ps->M * kHid, ps->M, kHid, cs, es, &L.step[d][buf], wn.c_str()));
            ps->adopt();
        }
    ps->flops += 2.0 * 2.0 * rows * kHid * kGates;
    {
        Step st;
        st.kind = Step::LSTM;
        st.layer = layer;
        st.name = p + ".recurrence";
        ps->steps.push_back(st);
    }
    DV_TRY(ps->alloc(&q, rows * Nout * sizeof(__half)));
    *out = reinterpret_cast<__half*>(q);
    EpiSpec es;
    es.out = *out, es.out_ld = Nout;
    return add_linear(e, ps, p + ".emb", L.y, rows, 2 * kHid, Nout, es);
}

int build_pass(Engine* e, CrnnModel* m, Pass* ps, int M, int H, int W) {
    if (H != 32 || (W % 4) || W < 16) return set_err(e, DV_ERR_ARG, "crnn: inputs must be 32 high and a multiple of 4 wide (got %dx%d)", H, W);
    ps->e = e;
    ps->M = M, ps->W = W, ps->T = W / 4;
    Tensor x0, t, u;
    DV_TRY(ps->tensor(&x0, M, 32, W, 8));
    ps->gray = x0.p;
    {
        Step st;
        st.kind = Step::GRAY;
        st.b = x0;
        ps->steps.push_back(st);
    }
    DV_TRY(add_conv(e, ps, "conv0", x0, 64, 3, 3, 1, &t, 32, W));
    DV_TRY(add_pool(ps, Step::POOL22, t, &u));
    DV_TRY(add_conv(e, ps, "conv1", u, 128, 3, 3, 1, &t, 16, W / 2));
    DV_TRY(add_pool(ps, Step::POOL22, t, &u));
    DV_TRY(add_conv(e, ps, "conv2a", u, 256, 3, 3, 1, &t, 8, W / 4));
    DV_TRY(add_conv(e, ps, "conv2b", t, 256, 3, 3, 1, &u, 8, W / 4));
    DV_TRY(add_pool(ps, Step::POOLH2, u, &t));
    DV_TRY(add_conv(e, ps, "conv3a", t, 512, 3, 3, 1, &u, 4, W / 4));
    DV_TRY(add_conv(e, ps, "conv3b", u, 512, 3, 3, 1, &t, 4, W / 4));
    DV_TRY(add_pool(ps, Step::POOLH2, t, &u));
    DV_TRY(add_conv(e, ps, "conv4", u, 512, 2, 1, 0, &t, 1, W / 4));  // [M,1,T,512] = the sequence [M*T, 512]
    __half *s1 = nullptr, *s2 = nullptr;
    DV_TRY(add_bilstm(e, ps, 0, t.p, 512, 256, &s1));
    DV_TRY(add_bilstm(e, ps, 1, s1, 256, 512, &s2));
    EpiSpec es;
    es.out = nullptr;  // logits dump, patched per call together with arg_out / max_out
    es.out_ld = m->labels;
    es.out_f32 = 1;
    es.arg_out = reinterpret_cast<int32_t*>(s1);  // placeholder
    DV_TRY(add_linear(e, ps, "cls", s2, static_cast<long long>(M) * ps->T, 512, m->labels, es, /*bias=*/false));
    ps->cls_step = static_cast<int>(ps->steps.size()) - 1;
    return 0;
}

int run_lstm(Engine* e, Pass* ps, int layer) {
    LstmLayer& L = ps->lstm[layer];
    const int M = ps->M, T = ps->T;
    DV_CUDA(e, cudaMemsetAsync(L.c, 0, static_cast<size_t>(2) * M * kHid * sizeof(float), e->stream));
    DV_CUDA(e, cudaMemsetAsync(L.h[0], 0, static_cast<size_t>(2) * M * kHid * sizeof(__half), e->stream));
    const long long total = 2LL * M * kHid;
    for (int t = 0; t < T; ++t) {
        const int buf = t & 1;
        for (int d = 0; d < 2; ++d) {
            ConvPlan& pl = L.step[d][buf];
            const int ts = d ? T - 1 - t : t;
            pl.prm.res = L.xproj + static_cast<size_t>(ts) * 2 * kGates + d * kGates;
            DV_TRY(launch_conv(e, pl));
        }
        e->launch_begin("k_lstm_cell", "lstm", 0.0, total * 30.0);
        k_lstm_cell<<<grid_for(total, 256), 256, 0, e->stream>>>(L.gates, L.c, L.h[buf ^ 1], L.y, M, T, t);
        e->launch_end();
    }
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

int run_pass(Engine* e, Pass* ps, const float* in, float* logits, int32_t* ids, float* maxv) {
    for (size_t i = 0; i < ps->steps.size(); ++i) {
        Step& st = ps->steps[i];
        switch (st.kind) {
            case Step::GRAY: {
                const long long plane = 32LL * ps->W, npix = plane * ps->M;
                e->launch_begin("k_crnn_gray", "pre", 0.0, npix * 28.0);
                k_crnn_gray<<<grid_for(npix, 256), 256, 0, e->stream>>>(in, npix, plane, ps->gray);
                e->launch_end();
                break;
            }
            case Step::CONV: DV_TRY(launch_conv(e, st.plan)); break;
            case Step::POOL22: DV_TRY(op_maxpool2x2(e, st.a, st.b)); break;
            case Step::POOLH2: {
                const long long total8 = static_cast<long long>(st.b.elems() / 8);
                e->launch_begin("k_maxpool_h2", "pool", 0.0, total8 * 48.0);
                k_maxpool_h2<<<grid_for(total8, 256), 256, 0, e->stream>>>(st.a.p, total8, st.a.W, st.a.C, st.b.p);
                e->launch_end();
                break;
            }
            case Step::LSTM: DV_TRY(run_lstm(e, ps, st.layer)); break;
            case Step::GEMM:
                if (static_cast<int>(i) == ps->cls_step) {
                    st.plan.prm.out = logits;
                    st.plan.prm.arg_out = ids;
                    st.plan.prm.max_out = maxv;
                }
                DV_TRY(launch_conv(e, st.plan));
                break;
        }
    }
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

}  // namespace

int crnn_create(Engine* e) {
    auto* m = new CrnnModel();
    e->model.reset(m);
    const BlobTensor* w = e->find("cls.w");
    if (!w || w->ndim != 2 || w->dims[1] != 512) return set_err(e, DV_ERR_WEIGHTS, "crnn: missing classifier weights");
    m->labels = static_cast<int>(w->dims[0]);
    if (m->labels % 4) return set_err(e, DV_ERR_UNSUPPORTED, "crnn: num_labels %% 4 != 0");
    if (const char* s = getenv("DV_CRNN_PASS")) {
        if (atoi(s) > 0) m->pass_n = atoi(s);
    }
    return 0;
}

int crnn_labels(Engine* e) {
    CrnnModel* m = dynamic_cast<CrnnModel*>(e->model.get());
    return m ? m->labels : 0;
}

double crnn_flops(Engine* e) {
    CrnnModel* m = dynamic_cast<CrnnModel*>(e->model.get());
    return m ? m->last_flops : 0.0;
}

// in fp32 [N,3,32,W] (the reference pre-processor's batch: RGB / 255) -> ids int32 [N, W/4] (+ logits fp32 [N, W/4, labels],
// + max logit [N, W/4])
int crnn_forward(Engine* e, const float* in, int N, int H, int W, float* logits, int32_t* ids, float* maxv) {
    CrnnModel* m = dynamic_cast<CrnnModel*>(e->model.get());
    if (!m) return set_err(e, DV_ERR_STATE, "handle was not created as a crnn model");
    if (N < 0 || (N > 0 && (!in || !ids))) return set_err(e, DV_ERR_ARG, "crnn_forward: bad arguments");
    m->last_flops = 0;
    const int T = W / 4;
    for (int done = 0; done < N;) {
        const int cur = (N - done) < m->pass_n ? (N - done) : m->pass_n;
        const long long key = static_cast<long long>(cur) * 4096 + W;
        auto it = m->passes.find(key);
        if (it == m->passes.end()) {
            if (m->passes.size() >= 4) m->passes.clear();  // bound the plan cache (each pass owns its buffers)
            std::unique_ptr<Pass> ps(new Pass());
            DV_TRY(build_pass(e, m, ps.get(), cur, H, W));
            it = m->passes.emplace(key, std::move(ps)).first;
        }
        Pass* ps = it->second.get();
        DV_TRY(run_pass(e, ps, in + static_cast<long long>(done) * 3 * H * W, logits ? logits + static_cast<long long>(done) * T * m->labels : nullptr,
                        ids + static_cast<long long>(done) * T, maxv ? maxv + static_cast<long long>(done) * T : nullptr));
        m->last_flops += ps->flops;
        done += cur;
    }
    return 0;
}

}  // namespace dv
