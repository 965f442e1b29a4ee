// Graph-described network executor: runs the op program that pdf_table_b200/picodet_graph.py lowers PicoDet to
// (LCNet-x1.0 + CSP-PAN + PicoHead; reference picodet/lcnet.py:159-263, picodet/csp_pan.py:233-360,
// picodet/pico_head.py:37-167, 1108-1138).  The CPU mirror is oracle/picodet_net_ref.py.
//
//   OP_STEM  uint8 HWC (or fp32 NCHW) image -> [normalise] -> 3x3 stride-2 conv 3 -> 16 + folded BN + hardswish
//            (OCRPicodetPreProcessor.normalize, picodet/processor_picodet.py:66-70, fused: fp32, numpy's operation order)
//   OP_DW    depthwise k x k (3 / 5), stride 1 / 2, folded BN, activation          CUDA cores, HBM-bound
//   OP_PW    1x1 conv + folded BN + activation                                     conv_igemm_tcgen05 (flat GEMM)
//   OP_SE    x * hardsigmoid(W2 relu(W1 avgpool(x) + b1) + b2)                     (LCNet SEModule :130-156)
//   OP_UP2   nearest 2x up-sampling into a concatenation slice                     (CSPPAN.forward :322-325)
//   OP_ADD   element-wise sum                                                      (the extra top level :343-345)
//   OP_HEAD  1x1 conv 128 -> C + 32 (fp32) -> sigmoid(class scores) [N,HW,C] + raw DFL logits [N,HW,32]
// Every tensor is NHWC fp16; an operand may be a channel slice (coff, c) of a wider buffer, which is how the CSP
// concatenations exist without copies.  The plan (buffers, TMA descriptors) is built once per input shape.
#include "engine.h"

namespace dv {

namespace {

enum { OP_STEM = 0, OP_DW, OP_PW, OP_SE, OP_UP2, OP_ADD, OP_HEAD };

__device__ __forceinline__ float act_f(float x, int act) {
    if (act == ACT_HSWISH) return x * fminf(fmaxf(x + 3.f, 0.f), 6.f) * (1.f / 6.f);
    if (act == ACT_RELU) return fmaxf(x, 0.f);
    return x;
}

// image -> 16 channels at half resolution.  One thread per output pixel; weights [27][16] (tap-major: r, s, c) in smem.
__global__ void __launch_bounds__(128)
k_stem3x3s2(const uint8_t* __restrict__ u8, const float* __restrict__ f32, int N, int H, int W, int Ho, int Wo, float3 mean,
            float3 stdv, float scale, int flip, const float* __restrict__ w, const float* __restrict__ bias, int act,
            __half* __restrict__ out) {
    __shared__ float sw[27 * 16];
    __shared__ float sb[16];
    for (int i = threadIdx.x; i < 27 * 16; i += blockDim.x) sw[i] = w[i];
    if (threadIdx.x < 16) sb[threadIdx.x] = bias[threadIdx.x];
    __syncthreads();
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<long long>(N) * Ho * Wo) return;
    const int ox = static_cast<int>(idx % Wo), oy = static_cast<int>((idx / Wo) % Ho), n = static_cast<int>(idx / (static_cast<long long>(Wo) * Ho));
    float acc[16];
#pragma unroll
    for (int c = 0; c < 16; ++c) acc[c] = sb[c];
    for (int r = 0; r < 3; ++r) {
        const int iy = 2 * oy - 1 + r;
        if (iy < 0 || iy >= H) continue;
        for (int s = 0; s < 3; ++s) {
            const int ix = 2 * ox - 1 + s;
            if (ix < 0 || ix >= W) continue;
            float v[3];
            if (u8) {
                const uint8_t* ip = u8 + ((static_cast<long long>(n) * H + iy) * W + ix) * 3;
                float c0 = ip[0], c1 = ip[1], c2 = ip[2];
                if (flip) {
                    const float t = c0;
                    c0 = c2;
                    c2 = t;
                }
                v[0] = __fdiv_rn(__fsub_rn(__fmul_rn(c0, scale), mean.x), stdv.x);
                v[1] = __fdiv_rn(__fsub_rn(__fmul_rn(c1, scale), mean.y), stdv.y);
                v[2] = __fdiv_rn(__fsub_rn(__fmul_rn(c2, scale), mean.z), stdv.z);
            } else {
                const long long plane = static_cast<long long>(H) * W;
                const float* ip = f32 + static_cast<long long>(n) * 3 * plane + static_cast<long long>(iy) * W + ix;
                v[0] = ip[0], v[1] = ip[plane], v[2] = ip[2 * plane];
            }
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                const float* wp = sw + ((r * 3 + s) * 3 + ci) * 16;
#pragma unroll
                for (int c = 0; c < 16; ++c) acc[c] = fmaf(v[ci], wp[c], acc[c]);
            }
        }
    }
    uint4 o[2];
    __half2* h = reinterpret_cast<__half2*>(o);
#pragma unroll
    for (int c = 0; c < 8; ++c) h[c] = __floats2half2_rn(act_f(acc[2 * c], act), act_f(acc[2 * c + 1], act));
    uint4* op = reinterpret_cast<uint4*>(out + idx * 16);
    op[0] = o[0];
    op[1] = o[1];
}

// depthwise k x k, pad (k-1)/2, stride s; w fp32 [k*k][C] (BN scale folded), b fp32 [C]; one thread = (pixel, 8 channels)
__global__ void __launch_bounds__(256)
k_dwconv(const __half* __restrict__ in, int N, int H, int W, int C, int ldi, int k, int stride, int Ho, int Wo,
         const float* __restrict__ w, const float* __restrict__ b, int act, __half* __restrict__ out, int ldo) {
    const int cv = C >> 3;
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<long long>(N) * Ho * Wo * cv) return;
    const int c8 = static_cast<int>(idx % cv);
    long long t = idx / cv;
    const int ox = static_cast<int>(t % Wo);
    t /= Wo;
    const int oy = static_cast<int>(t % Ho), n = static_cast<int>(t / Ho);
    float acc[8];
    {
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + c8 * 8)), b1 = __ldg(reinterpret_cast<const float4*>(b + c8 * 8 + 4));
        acc[0] = b0.x, acc[1] = b0.y, acc[2] = b0.z, acc[3] = b0.w, acc[4] = b1.x, acc[5] = b1.y, acc[6] = b1.z, acc[7] = b1.w;
    }
    const int pad = (k - 1) >> 1;
    for (int r = 0; r < k; ++r) {
        const int iy = oy * stride - pad + r;
        if (iy < 0 || iy >= H) continue;
        for (int s = 0; s < k; ++s) {
            const int ix = ox * stride - pad + s;
            if (ix < 0 || ix >= W) continue;
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(in + ((static_cast<long long>(n) * H + iy) * W + ix) * ldi + c8 * 8));
            const float4* wp = reinterpret_cast<const float4*>(w + static_cast<long long>(r * k + s) * C + c8 * 8);
            const float4 w0 = __ldg(wp), w1 = __ldg(wp + 1);
            const __half2* h = reinterpret_cast<const __half2*>(&u);
            const float2 v0 = __half22float2(h[0]), v1 = __half22float2(h[1]), v2 = __half22float2(h[2]), v3 = __half22float2(h[3]);
            acc[0] = fmaf(v0.x, w0.x, acc[0]);
            acc[1] = fmaf(v0.y, w0.y, acc[1]);
            acc[2] = fmaf(v1.x, w0.z, acc[2]);
            acc[3] = fmaf(v1.y, w0.w, acc[3]);
            acc[4] = fmaf(v2.x, w1.x, acc[4]);
            acc[5] = fmaf(v2.y, w1.y, acc[5]);
            acc[6] = fmaf(v3.x, w1.z, acc[6]);
            acc[7] = fmaf(v3.y, w1.w, acc[7]);
        }
    }
    uint4 o;
    __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) ho[i] = __floats2half2_rn(act_f(acc[2 * i], act), act_f(acc[2 * i + 1], act));
    *reinterpret_cast<uint4*>(out + ((static_cast<long long>(n) * Ho + oy) * Wo + ox) * ldo + c8 * 8) = o;
}

// SE squeeze: one CTA per image.  avg[c] -> hidden = relu(W1 avg + b1) -> scale[c] = hardsigmoid(W2 hidden + b2)
__global__ void __launch_bounds__(512)
k_se_scale(const __half* __restrict__ in, int HW, int C, const float* __restrict__ w1, const float* __restrict__ b1,
           const float* __restrict__ w2, const float* __restrict__ b2, float* __restrict__ scale) {
    extern __shared__ float sm[];  // avg[C] | hidden[C/4]
    float* avg = sm;
    float* hid = sm + C;
    const int n = blockIdx.x, R = C >> 2;
    const __half* base = in + static_cast<long long>(n) * HW * C;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = 0.f;
        for (int p = 0; p < HW; ++p) s += __half2float(base[static_cast<long long>(p) * C + c]);
        avg[c] = s / static_cast<float>(HW);
    }
    __syncthreads();
    for (int r = threadIdx.x; r < R; r += blockDim.x) {
        float s = b1[r];
        for (int c = 0; c < C; ++c) s = fmaf(w1[r * C + c], avg[c], s);
        hid[r] = fmaxf(s, 0.f);
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float s = b2[c];
        for (int r = 0; r < R; ++r) s = fmaf(w2[c * R + r], hid[r], s);
        scale[n * C + c] = fminf(fmaxf(s * (1.f / 6.f) + 0.5f, 0.f), 1.f);  // F.hardsigmoid
    }
}

__global__ void __launch_bounds__(256)
k_se_apply(const __half* __restrict__ in, long long total8, int HW, int C, const float* __restrict__ scale, __half* __restrict__ out) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= total8) return;
    const int cv = C >> 3;
    const int c8 = static_cast<int>(idx % cv);
    const int n = static_cast<int>(idx / (static_cast<long long>(cv) * HW));
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(in) + idx);
    const float* sp = scale + n * C + c8 * 8;
    const __half2* h = reinterpret_cast<const __half2*>(&u);
    uint4 o;
    __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 v = __half22float2(h[i]);
        ho[i] = __floats2half2_rn(v.x * sp[2 * i], v.y * sp[2 * i + 1]);
    }
    reinterpret_cast<uint4*>(out)[idx] = o;
}

__global__ void __launch_bounds__(256)
k_up2(const __half* __restrict__ in, int N, int h, int w, int C, int ldi, int Ho, int Wo, __half* __restrict__ out, int ldo) {
    const int cv = C >> 3;
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= static_cast<long long>(N) * Ho * Wo * cv) return;
    const int c8 = static_cast<int>(idx % cv);
    long long t = idx / cv;
    const int ox = static_cast<int>(t % Wo);
    t /= Wo;
    const int oy = static_cast<int>(t % Ho), n = static_cast<int>(t / Ho);
    // F.interpolate(mode="nearest") to an explicit size: src = floor(dst * in / out)
    const int iy = min(static_cast<int>(static_cast<long long>(oy) * h / Ho), h - 1), ix = min(static_cast<int>(static_cast<long long>(ox) * w / Wo), w - 1);
    const uint4 u = __ldg(reinterpret_cast<const uint4*>(in + ((static_cast<long long>(n) * h + iy) * w + ix) * ldi + c8 * 8));
    *reinterpret_cast<uint4*>(out + ((static_cast<long long>(n) * Ho + oy) * Wo + ox) * ldo + c8 * 8) = u;
}

__global__ void __launch_bounds__(256) k_add(const __half* __restrict__ a, const __half* __restrict__ b, long long total8, __half* __restrict__ out) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= total8) return;
    const uint4 ua = __ldg(reinterpret_cast<const uint4*>(a) + idx), ub = __ldg(reinterpret_cast<const uint4*>(b) + idx);
    const __half2 *ha = reinterpret_cast<const __half2*>(&ua), *hb = reinterpret_cast<const __half2*>(&ub);
    uint4 o;
    __half2* ho = reinterpret_cast<__half2*>(&o);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float2 x = __half22float2(ha[i]), y = __half22float2(hb[i]);
        ho[i] = __floats2half2_rn(x.x + y.x, x.y + y.y);
    }
    reinterpret_cast<uint4*>(out)[idx] = o;
}

// raw fp32 [M, ld] head output -> scores [M, C] = sigmoid(raw[:, :C]), dfl [M, R] = raw[:, C:C+R]
__global__ void __launch_bounds__(256)
k_head_split(const float* __restrict__ raw, long long M, int ld, int C, int R, float* __restrict__ scores, float* __restrict__ dfl) {
    const long long idx = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x;
    if (idx >= M * (C + R)) return;
    const long long m = idx / (C + R);
    const int j = static_cast<int>(idx % (C + R));
    const float v = raw[m * ld + j];
    if (j < C) scores[m * C + j] = 1.f / (1.f + expf(-v));
    else dfl[m * R + (j - C)] = v;
}

struct GOp {
    int code, in_t, in_coff, in_c, out_t, out_coff, out_c, k, stride, act, w, aux;
    ConvPlan plan;  // OP_PW / OP_HEAD
    const float *f0 = nullptr, *f1 = nullptr, *f2 = nullptr, *f3 = nullptr;
};

struct GraphNet : Model {
    Engine* e = nullptr;
    int N = 0, H = 0, W = 0;
    int num_classes = 0, reg_bins = 32, head_ld = 40;
    std::vector<int> tc, tdown;
    std::vector<GOp> ops;
    std::vector<Tensor> tens;
    std::vector<void*> mem;
    float* se_scale = nullptr;
    float* head_raw = nullptr;
    double flops = 0;
    ~GraphNet() override {
        for (void* p : mem) cudaFree(p);
    }
    int alloc(void** p, size_t bytes) {
        cudaError_t st = cudaMalloc(p, bytes ? bytes : 16);
        if (st != cudaSuccess) return set_err(e, DV_ERR_CUDA, "graph: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(st));
        mem.push_back(*p);
        return 0;
    }
};

const float* wf32(Engine* e, int id, const char* field, size_t elems, int* rc) {
    const std::string name = "w" + std::to_string(id) + "." + field;
    const BlobTensor* t = e->find(name);
    if (!t || t->dtype != 0 || t->nbytes < elems * 4) {
        *rc = set_err(e, DV_ERR_WEIGHTS, "graph: missing / short fp32 tensor '%s'", name.c_str());
        return nullptr;
    }
    return reinterpret_cast<const float*>(t->dptr);
}

int build(Engine* e, GraphNet* m, int N, int H, int W) {
    for (void* p : m->mem) cudaFree(p);
    m->mem.clear();
    m->e = e;
    m->N = N;
    m->H = H;
    m->W = W;
    m->flops = 0;
    const size_t nt = m->tc.size();
    m->tens.assign(nt, Tensor());
    for (size_t i = 1; i < nt; ++i) {  // tensor 0 is the input image
        Tensor& t = m->tens[i];
        t.N = N;
        t.H = (H + m->tdown[i] - 1) / m->tdown[i];
        t.W = (W + m->tdown[i] - 1) / m->tdown[i];
        t.C = m->tc[i];
        void* p = nullptr;
        DV_TRY(m->alloc(&p, t.elems() * sizeof(__half)));
        t.p = reinterpret_cast<__half*>(p);
    }
    size_t max_head_rows = 0;
    int max_se_c = 0;
    for (GOp& op : m->ops) {
        int rc = 0;
        const Tensor& in = m->tens[op.in_t];
        const Tensor& out = m->tens[op.out_t];
        switch (op.code) {
            case OP_STEM:
                op.f0 = wf32(e, op.w, "sw", 27 * 16, &rc);
                op.f1 = wf32(e, op.w, "sb", 16, &rc);
                m->flops += 2.0 * N * out.H * out.W * 27 * 16;
                break;
            case OP_DW:
                op.f0 = wf32(e, op.w, "dw", static_cast<size_t>(op.k) * op.k * op.in_c, &rc);
                op.f1 = wf32(e, op.w, "db", op.in_c, &rc);
                m->flops += 2.0 * N * out.H * out.W * op.k * op.k * op.in_c;
                break;
            case OP_SE:
                op.f0 = wf32(e, op.w, "s1w", static_cast<size_t>(op.in_c) * op.in_c / 4, &rc);
                op.f1 = wf32(e, op.w, "s1b", op.in_c / 4, &rc);
                op.f2 = wf32(e, op.w, "s2w", static_cast<size_t>(op.in_c) * op.in_c / 4, &rc);
                op.f3 = wf32(e, op.w, "s2b", op.in_c, &rc);
                if (op.in_c > max_se_c) max_se_c = op.in_c;
                break;
            case OP_HEAD:
                if (static_cast<size_t>(N) * in.H * in.W > max_head_rows) max_head_rows = static_cast<size_t>(N) * in.H * in.W;
                break;
            default: break;
        }
        if (rc) return rc;
    }
    if (max_se_c) {
        void* p = nullptr;
        DV_TRY(m->alloc(&p, static_cast<size_t>(N) * max_se_c * 4));
        m->se_scale = reinterpret_cast<float*>(p);
    }
    {
        void* p = nullptr;
        DV_TRY(m->alloc(&p, max_head_rows * m->head_ld * 4 + 16));
        m->head_raw = reinterpret_cast<float*>(p);
    }
    for (GOp& op : m->ops) {
        if (op.code != OP_PW && op.code != OP_HEAD) continue;
        const Tensor& in = m->tens[op.in_t];
        const Tensor& out = m->tens[op.out_t];
        const std::string wn = "w" + std::to_string(op.w);
        const BlobTensor* w = e->find(wn + ".w");
        const BlobTensor* b = e->find(wn + ".b");
        if (!w || !b || w->dtype != 1 || b->dtype != 0 || w->ndim != 2 || static_cast<int>(w->dims[0]) != op.out_c ||
            static_cast<int>(w->dims[1]) != op.in_c)
            return set_err(e, DV_ERR_WEIGHTS, "graph: bad 1x1 weights '%s' (want [%d,%d])", wn.c_str(), op.out_c, op.in_c);
        ConvSpec cs;
        cs.KH = cs.KW = 1;
        cs.Cin = op.in_c;
        cs.Cin_pad = op.in_c;
        cs.Cout = op.out_c;
        cs.BK = (op.in_c % 64 == 0) ? 64 : (op.in_c % 32 == 0) ? 32 : 16;
        cs.w = reinterpret_cast<const __half*>(w->dptr);
        cs.bias = reinterpret_cast<const float*>(b->dptr);
        cs.flat = true;
        EpiSpec es;
        es.act = op.act;
        if (op.code == OP_HEAD) {
            es.out = m->head_raw;
            es.out_ld = m->head_ld;
            es.out_f32 = 1;
        } else {
            es.out = out.p;
            es.out_ld = out.C;
            es.out_coff = op.out_coff;
        }
        const int M = N * in.H * in.W;
        DV_TRY(plan_linear(e, in.p + op.in_coff, M, op.in_c, cs, es, &op.plan, wn.c_str(), in.C));
        m->mem.push_back(e->owned.back());
        e->owned.pop_back();
        m->flops += op.plan.flops;
    }
    return 0;
}

static inline int grid_for(long long n, int block) { return static_cast<int>((n + block - 1) / block); }

}  // namespace

int graph_create(Engine* e) {
    GraphNet* m = new GraphNet();
    m->e = e;
    const BlobTensor* tt = e->find("graph.tensors");
    const BlobTensor* to = e->find("graph.ops");
    const BlobTensor* tm = e->find("graph.meta");
    if (!tt || !to || !tm || tt->dtype != 2 || to->dtype != 2 || tm->dtype != 2 || to->dims[1] != 12 || tt->dims[1] != 2) {
        delete m;
        return set_err(e, DV_ERR_WEIGHTS, "graph model: missing graph.tensors / graph.ops / graph.meta");
    }
    std::vector<int32_t> ht(tt->nbytes / 4), ho(to->nbytes / 4), hm(tm->nbytes / 4);
    cudaMemcpy(ht.data(), tt->dptr, tt->nbytes, cudaMemcpyDeviceToHost);
    cudaMemcpy(ho.data(), to->dptr, to->nbytes, cudaMemcpyDeviceToHost);
    cudaMemcpy(hm.data(), tm->dptr, tm->nbytes, cudaMemcpyDeviceToHost);
    m->num_classes = hm[0];
    m->reg_bins = hm[1];
    m->head_ld = hm[2];
    for (size_t i = 0; i < tt->dims[0]; ++i) {
        m->tc.push_back(ht[2 * i]);
        m->tdown.push_back(ht[2 * i + 1]);
    }
    for (size_t i = 0; i < to->dims[0]; ++i) {
        const int32_t* o = &ho[12 * i];
        GOp op;
        op.code = o[0], op.in_t = o[1], op.in_coff = o[2], op.in_c = o[3], op.out_t = o[4], op.out_coff = o[5], op.out_c = o[6];
        op.k = o[7], op.stride = o[8], op.act = o[9], op.w = o[10], op.aux = o[11];
        const int nt = static_cast<int>(m->tc.size());
        if (op.in_t < 0 || op.in_t >= nt || op.out_t < 0 || op.out_t >= nt || (op.in_coff % 8) || (op.out_coff % 8) || (op.in_c % 8 && op.code != OP_STEM)) {
            delete m;
            return set_err(e, DV_ERR_WEIGHTS, "graph model: malformed op %zu", i);
        }
        m->ops.push_back(op);
    }
    e->model.reset(m);
    return 0;
}

double graph_flops(Engine* e) {
    GraphNet* m = dynamic_cast<GraphNet*>(e->model.get());
    return m ? m->flops : 0.0;
}

int graph_num_classes(Engine* e) {
    GraphNet* m = dynamic_cast<GraphNet*>(e->model.get());
    return m ? m->num_classes : 0;
}

int graph_debug_tensor(Engine* e, int tensor_id, float* out_nchw, int* dims4) {
    GraphNet* m = dynamic_cast<GraphNet*>(e->model.get());
    if (!m) return set_err(e, DV_ERR_STATE, "not a graph-model handle");
    if (tensor_id <= 0 || tensor_id >= static_cast<int>(m->tens.size()) || !m->tens[tensor_id].p) return set_err(e, DV_ERR_ARG, "no tensor %d", tensor_id);
    const Tensor& t = m->tens[tensor_id];
    if (dims4) {
        dims4[0] = t.N, dims4[1] = t.C, dims4[2] = t.H, dims4[3] = t.W;
    }
    if (out_nchw) return op_nhwc_f16_to_nchw_f32(e, t.p, t.N, t.C, t.H, t.W, out_nchw);
    return 0;
}

// scores_out[l] fp32 [N, HW_l, C], dfl_out[l] fp32 [N, HW_l, 32] (device), l = 0..3
int picodet_forward(Engine* e, const float* in_nchw, const uint8_t* in_u8, const float* mean3, const float* std3, float scale, int flip,
                    int N, int H, int W, float* const* scores_out, float* const* dfl_out) {
    GraphNet* m = dynamic_cast<GraphNet*>(e->model.get());
    if (!m) return set_err(e, DV_ERR_STATE, "handle was not created as a picodet model");
    if (N <= 0 || H <= 0 || W <= 0 || (!in_nchw && !in_u8) || !scores_out || !dfl_out) return set_err(e, DV_ERR_ARG, "picodet_forward: bad arguments");
    if (m->N != N || m->H != H || m->W != W) DV_TRY(build(e, m, N, H, W));
    cudaStream_t s = e->stream;
    for (GOp& op : m->ops) {
        const Tensor& in = m->tens[op.in_t];
        const Tensor& out = m->tens[op.out_t];
        switch (op.code) {
            case OP_STEM: {
                const long long total = static_cast<long long>(N) * out.H * out.W;
                float3 mean = make_float3(0, 0, 0), stdv = make_float3(1, 1, 1);
                if (in_u8) {
                    mean = make_float3(mean3[0], mean3[1], mean3[2]);
                    stdv = make_float3(std3[0], std3[1], std3[2]);
                }
                e->launch_begin("k_stem3x3s2", "conv1", 2.0 * total * 27 * 16, total * (12.0 * (in_u8 ? 1 : 4) + 32.0));
                k_stem3x3s2<<<grid_for(total, 128), 128, 0, s>>>(in_u8, in_nchw, N, H, W, out.H, out.W, mean, stdv, scale, flip, op.f0, op.f1, op.act,
                                                                 out.p);
                e->launch_end();
                break;
            }
            case OP_DW: {
                const long long total = static_cast<long long>(N) * out.H * out.W * (op.in_c / 8);
                e->launch_begin("k_dwconv", "dw", 2.0 * total * 8 * op.k * op.k, total * 8 * 2.0 * (1.0 + 1.0 / (op.stride * op.stride)));
                k_dwconv<<<grid_for(total, 256), 256, 0, s>>>(in.p + op.in_coff, N, in.H, in.W, op.in_c, in.C, op.k, op.stride, out.H, out.W, op.f0, op.f1,
                                                              op.act, out.p + op.out_coff, out.C);
                e->launch_end();
                break;
            }
            case OP_PW: DV_TRY(launch_conv(e, op.plan)); break;
            case OP_SE: {
                const int HW = in.H * in.W;
                e->launch_begin("k_se_scale", "se", 0.0, static_cast<double>(N) * HW * op.in_c * 2.0);
                k_se_scale<<<N, 512, (op.in_c + op.in_c / 4) * sizeof(float), s>>>(in.p, HW, op.in_c, op.f0, op.f1, op.f2, op.f3, m->se_scale);
                e->launch_end();
                const long long total8 = static_cast<long long>(N) * HW * (op.in_c / 8);
                e->launch_begin("k_se_apply", "se", 0.0, total8 * 32.0);
                k_se_apply<<<grid_for(total8, 256), 256, 0, s>>>(in.p, total8, HW, op.in_c, m->se_scale, out.p);
                e->launch_end();
                break;
            }
            case OP_UP2: {
                const long long total = static_cast<long long>(N) * out.H * out.W * (op.in_c / 8);
                e->launch_begin("k_up2", "up", 0.0, total * 16.0 * 1.25);
                k_up2<<<grid_for(total, 256), 256, 0, s>>>(in.p + op.in_coff, N, in.H, in.W, op.in_c, in.C, out.H, out.W, out.p + op.out_coff, out.C);
                e->launch_end();
                break;
            }
            case OP_ADD: {
                const Tensor& b = m->tens[op.aux];
                const long long total8 = static_cast<long long>(out.elems() / 8);
                e->launch_begin("k_add", "add", 0.0, total8 * 48.0);
                k_add<<<grid_for(total8, 256), 256, 0, s>>>(in.p, b.p, total8, out.p);
                e->launch_end();
                break;
            }
            case OP_HEAD: {
                DV_TRY(launch_conv(e, op.plan));
                const long long M = static_cast<long long>(N) * in.H * in.W;
                if (op.aux < 0 || op.aux > 3 || !scores_out[op.aux] || !dfl_out[op.aux]) return set_err(e, DV_ERR_ARG, "picodet_forward: null output for level %d", op.aux);
                e->launch_begin("k_head_split", "head", 0.0, M * (m->num_classes + m->reg_bins) * 8.0);
                k_head_split<<<grid_for(M * (m->num_classes + m->reg_bins), 256), 256, 0, s>>>(m->head_raw, M, m->head_ld, m->num_classes, m->reg_bins,
                                                                                               scores_out[op.aux], dfl_out[op.aux]);
                e->launch_end();
                break;
            }
            default: return set_err(e, DV_ERR_WEIGHTS, "graph model: unknown opcode %d", op.code);
        }
    }
    DV_CUDA(e, cudaGetLastError());
    return 0;
}

}  // namespace dv
