// DBNet (ResNet-18 backbone + FPN SegDetector head) as a static plan of conv_igemm_tcgen05 launches.
// Architecture restated from the reference: ResNet db_net/dbnet.py:260-335 (BasicBlock :103-169),
// SegDetector.forward :618-650 (eval branch returns `binary`), DBModel :715-728.
// BatchNorm is folded into the fp16 weights / fp32 bias by pdf_table_b200/weights.py.
#include "engine.h"

namespace dv {

namespace {

struct Step {
    enum Kind { CONV, MAXPOOL, DECONV_FINAL } kind;
    ConvPlan plan;
    Tensor a, b;
};

// fp32x mode (blob entry "precision", weights.pack_dbnet_r18(precise=True)): every activation tensor is a split-fp16 pair
// ([hi(C) | lo(C)] per pixel; the padded stem image as two image batches), every weight a [W_hi | W_lo | W_hi] triple per
// filter tap, and conv_igemm_tcgen05 walks (hi, hi, lo) k-blocks per tap: three MMAs per product, fp32 TMEM accumulation.
// The probability map lands within 1e-3 of the fp32 oracle (tests/test_gpu_dbnet.py) where the fp16-operand default is 2.4e-3.
struct DbNet : Model {
    bool precise = false;
    const float* final_w32 = nullptr;
    int N = 0, H = 0, W = 0;
    std::vector<Step> steps;
    Tensor stem_in;  // padded [N, H+6, W+8, 4]
    Tensor last;     // [N, H/2, W/2, 64] input of the final deconv
    const __half* final_w = nullptr;
    float final_b = 0.f;
    std::map<std::string, Tensor> named;  // intermediate tensors by name (dv_debug_get_tensor)
    double flops = 0;
};

int get_conv(Engine* e, const std::string& name, ConvSpec* cs) {
    const BlobTensor* w = e->find(name + ".w");
    const BlobTensor* b = e->find(name + ".b");
    if (!w || !b) return set_err(e, DV_ERR_WEIGHTS, "missing weights for '%s'", name.c_str());
    if (w->dtype != 1 || b->dtype != 0 || w->ndim != 2)
        return set_err(e, DV_ERR_WEIGHTS, "bad dtype/rank for '%s'", name.c_str());
    cs->w = reinterpret_cast<const __half*>(w->dptr);
    cs->bias = reinterpret_cast<const float*>(b->dptr);
    const int parts = cs->split ? 3 : 1;
    const int taps = cs->KH * cs->KW * parts;
    if (cs->stem) {
        if (w->dims[0] != 64 || static_cast<int>(w->dims[1]) != 224 * parts)
            return set_err(e, DV_ERR_WEIGHTS, "'%s': stem weight must be [64,%d]", name.c_str(), 224 * parts);
    } else if (static_cast<int>(w->dims[0]) != cs->Cout || (w->dims[1] % taps) != 0)
        return set_err(e, DV_ERR_WEIGHTS, "'%s': weight shape [%u,%u] does not match Cout=%d taps=%d", name.c_str(),
                       w->dims[0], w->dims[1], cs->Cout, taps);
    cs->Cin_pad = cs->stem ? 32 : static_cast<int>(w->dims[1]) / taps;
    if (cs->split && !cs->stem && cs->KH == 1 && cs->stride == 1) cs->Cin_pad *= 3;  // flat GEMM: plan_linear counts all three parts
    if (!cs->stem) cs->BK = (cs->Cin_pad % 64 == 0) ? 64 : (cs->Cin_pad % 32 == 0) ? 32 : 16;
    if (b->dims[0] < static_cast<uint32_t>((cs->Cout + 255) / 256 * 256))
        return set_err(e, DV_ERR_WEIGHTS, "'%s': bias not padded to 256", name.c_str());
    return 0;
}

int alloc_tensor(Engine* e, DbNet* m, Tensor* t, int N, int H, int W, int C, bool zero = false) {
    t->N = N;
    t->H = H;
    t->W = W;
    t->C = C;
    void* p = nullptr;
    DV_TRY(e->dalloc(&p, t->elems() * sizeof(__half) * (m->precise ? 2 : 1), zero));
    t->p = reinterpret_cast<__half*>(p);
    if (m->precise) {  // [hi(C) | lo(C)] per pixel
        t->ld = 2 * C;
        t->lo = C;
    }
    return 0;
}
#define NAMED(m, name, t) (m)->named[(name)] = (t)

int add_conv(Engine* e, DbNet* m, const std::string& name, const Tensor& in, int cout, int k, int stride, int pad,
             const EpiSpec& es, int Ho, int Wo, bool stem = false) {
    ConvSpec cs;
    cs.KH = cs.KW = k;
    cs.stride = stride;
    cs.pad = pad;
    cs.Cin = stem ? 3 : in.C;
    cs.Cout = cout;
    cs.stem = stem;
    cs.split = m->precise;
    if (stem) cs.BK = 32;
    DV_TRY(get_conv(e, name, &cs));
    Step st;
    st.kind = Step::CONV;
    DV_TRY(plan_conv(e, in, cs, es, Ho, Wo, &st.plan, name.c_str()));
    m->flops += st.plan.flops;
    m->steps.push_back(st);
    return 0;
}

EpiSpec epi(Tensor& out, int act, const Tensor* res = nullptr, int res_mode = RES_NONE) {
    EpiSpec es;
    es.out = out.p;
    es.out_ld = out.ldc();
    es.split_off = static_cast<int>(out.lo);
    es.act = act;
    if (res) {
        es.res = res->p;
        es.res_mode = res_mode;
        es.res_ld = res->ldc();
        es.res_lo = static_cast<int>(res->lo);
    }
    return es;
}

int build(Engine* e, DbNet* m, int N, int H, int W) {
    if ((H % 32) || (W % 32)) return set_err(e, DV_ERR_ARG, "dbnet: H and W must be multiples of 32 (got %dx%d)", H, W);
    m->N = N;
    m->H = H;
    m->W = W;
    m->steps.clear();
    m->named.clear();
    m->flops = 0;
    {  // the padded stem image is never a [hi | lo] pixel pair: fp32x keeps the lo copy as a second image batch behind it
        const bool pr = m->precise;
        m->precise = false;
        int rc = alloc_tensor(e, m, &m->stem_in, pr ? 2 * N : N, H + 6, W + 8, 4, /*zero=*/true);
        m->precise = pr;
        if (rc) return rc;
        if (pr) {
            m->stem_in.N = N;
            m->stem_in.lo = static_cast<long long>(N) * (H + 6) * (W + 8) * 4;
        }
    }
    Tensor c1, p1;
    DV_TRY(alloc_tensor(e, m, &c1, N, H / 2, W / 2, 64));
    DV_TRY(alloc_tensor(e, m, &p1, N, H / 4, W / 4, 64));
    DV_TRY(add_conv(e, m, "stem", m->stem_in, 64, 7, 2, 3, epi(c1, ACT_RELU), H / 2, W / 2, true));
    {
        Step st;
        st.kind = Step::MAXPOOL;
        st.a = c1;
        st.b = p1;
        m->steps.push_back(st);
    }
    NAMED(m, "c1", c1);
    NAMED(m, "p1", p1);
    // ---- residual stages
    Tensor x = p1;
    Tensor feats[4];
    const int planes[4] = {64, 128, 256, 512};
    for (int L = 0; L < 4; ++L) {
        for (int B = 0; B < 2; ++B) {
            const int stride = (L > 0 && B == 0) ? 2 : 1;
            const int Ho = x.H / stride, Wo = x.W / stride;
            const std::string pre = "layer" + std::to_string(L + 1) + "." + std::to_string(B);
            Tensor t, y, ds;
            DV_TRY(alloc_tensor(e, m, &t, N, Ho, Wo, planes[L]));
            DV_TRY(alloc_tensor(e, m, &y, N, Ho, Wo, planes[L]));
            DV_TRY(add_conv(e, m, pre + ".conv1", x, planes[L], 3, stride, 1, epi(t, ACT_RELU), Ho, Wo));
            const Tensor* res = &x;
            if (stride != 1 || x.C != planes[L]) {
                DV_TRY(alloc_tensor(e, m, &ds, N, Ho, Wo, planes[L]));
                DV_TRY(add_conv(e, m, pre + ".down", x, planes[L], 1, stride, 0, epi(ds, ACT_NONE), Ho, Wo));
                res = &ds;
            }
            DV_TRY(add_conv(e, m, pre + ".conv2", t, planes[L], 3, 1, 1, epi(y, ACT_RELU, res, RES_SAME), Ho, Wo));
            NAMED(m, pre + ".t", t);
            NAMED(m, pre, y);
            x = y;
        }
        feats[L] = x;
    }
    // ---- FPN (SegDetector.forward): in5..in2 lateral 1x1, top-down nearest-2x add fused in the epilogue
    Tensor in5, out4, out3, out2, fuse;
    DV_TRY(alloc_tensor(e, m, &in5, N, feats[3].H, feats[3].W, 256));
    DV_TRY(alloc_tensor(e, m, &out4, N, feats[2].H, feats[2].W, 256));
    DV_TRY(alloc_tensor(e, m, &out3, N, feats[1].H, feats[1].W, 256));
    DV_TRY(alloc_tensor(e, m, &out2, N, feats[0].H, feats[0].W, 256));
    DV_TRY(alloc_tensor(e, m, &fuse, N, feats[0].H, feats[0].W, 256));
    DV_TRY(add_conv(e, m, "in5", feats[3], 256, 1, 1, 0, epi(in5, ACT_NONE), in5.H, in5.W));
    DV_TRY(add_conv(e, m, "in4", feats[2], 256, 1, 1, 0, epi(out4, ACT_NONE, &in5, RES_UP2), out4.H, out4.W));
    DV_TRY(add_conv(e, m, "in3", feats[1], 256, 1, 1, 0, epi(out3, ACT_NONE, &out4, RES_UP2), out3.H, out3.W));
    DV_TRY(add_conv(e, m, "in2", feats[0], 256, 1, 1, 0, epi(out2, ACT_NONE, &out3, RES_UP2), out2.H, out2.W));
    // p5..p2: 3x3 256->64, nearest-upsampled by 8/4/2/1 straight into the channel slices of `fuse`
    // (torch.cat((p5, p4, p3, p2), 1), dbnet.py:633).
    {
        const Tensor* src[4] = {&in5, &out4, &out3, &out2};
        const char* names[4] = {"out5", "out4", "out3", "out2"};
        const int rep[4] = {8, 4, 2, 1};
        for (int i = 0; i < 4; ++i) {
            EpiSpec es;
            es.out = fuse.p;
            es.out_ld = fuse.ldc();
            es.split_off = static_cast<int>(fuse.lo);
            es.out_coff = 64 * i;
            es.out_mode = rep[i] > 1 ? OUT_REPL : OUT_NHWC;
            es.rep = rep[i];
            DV_TRY(add_conv(e, m, names[i], *src[i], 64, 3, 1, 1, es, src[i]->H, src[i]->W));
        }
    }
    NAMED(m, "in5", in5);
    NAMED(m, "out4", out4);
    NAMED(m, "out3", out3);
    NAMED(m, "out2", out2);
    NAMED(m, "fuse", fuse);
    // ---- binarize head: conv3x3+BN+ReLU -> ConvT2x2(64->64)+BN+ReLU (GEMM + pixel shuffle) -> ConvT2x2(64->1)+sigmoid
    Tensor b1, b2;
    DV_TRY(alloc_tensor(e, m, &b1, N, fuse.H, fuse.W, 64));
    DV_TRY(alloc_tensor(e, m, &b2, N, fuse.H * 2, fuse.W * 2, 64));
    DV_TRY(add_conv(e, m, "bin.conv", fuse, 64, 3, 1, 1, epi(b1, ACT_RELU), b1.H, b1.W));
    {
        EpiSpec es;
        es.out = b2.p;
        es.out_ld = b2.ldc();
        es.split_off = static_cast<int>(b2.lo);
        es.out_mode = OUT_SHUF2;
        es.act = ACT_RELU;
        DV_TRY(add_conv(e, m, "bin.deconv1", b1, 256, 1, 1, 0, es, b1.H, b1.W));
    }
    {
        const BlobTensor* w = e->find("bin.deconv2.w");
        const BlobTensor* b = e->find("bin.deconv2.b");
        if (!w || !b || w->dtype != 1 || b->dtype != 0)
            return set_err(e, DV_ERR_WEIGHTS, "missing bin.deconv2 weights");
        NAMED(m, "b1", b1);
        NAMED(m, "b2", b2);
        m->final_w = reinterpret_cast<const __half*>(w->dptr);
        if (m->precise) {
            const BlobTensor* w32 = e->find("bin.deconv2.w32");
            if (!w32 || w32->dtype != 0 || w32->nbytes < 256 * 4) return set_err(e, DV_ERR_WEIGHTS, "missing bin.deconv2.w32 (fp32x blob)");
            m->final_w32 = reinterpret_cast<const float*>(w32->dptr);
        }
        DV_CUDA(e, cudaMemcpy(&m->final_b, b->dptr, sizeof(float), cudaMemcpyDeviceToHost));
        m->last = b2;
        Step st;
        st.kind = Step::DECONV_FINAL;
        st.a = b2;
        m->steps.push_back(st);
        m->flops += 2.0 * N * b2.H * b2.W * 64 * 4;
    }
    return 0;
}

}  // namespace

int dbnet_create(Engine* e) {
    DbNet* m = new DbNet();
    m->precise = e->find("precision") != nullptr;
    e->model.reset(m);
    return 0;
}

int dbnet_debug_tensor(Engine* e, const char* name, float* out_nchw, int* dims4) {
    DbNet* m = dynamic_cast<DbNet*>(e->model.get());
    if (!m) return set_err(e, DV_ERR_STATE, "not a dbnet handle");
    auto it = m->named.find(name);
    if (it == m->named.end()) return set_err(e, DV_ERR_ARG, "no tensor named '%s'", name);
    const Tensor& t = it->second;
    if (dims4) {
        dims4[0] = t.N;
        dims4[1] = t.C;
        dims4[2] = t.H;
        dims4[3] = t.W;
    }
    if (out_nchw) return op_nhwc_f16_to_nchw_f32(e, t.p, t.N, t.C, t.H, t.W, out_nchw, t.ldc(), static_cast<int>(t.lo));
    return 0;
}

double dbnet_flops(Engine* e) {
    DbNet* m = dynamic_cast<DbNet*>(e->model.get());
    return m ? m->flops : 0.0;
}

int dbnet_forward(Engine* e, const float* in_nchw, const uint8_t* in_u8, const float* mean3, const float* std3,
                  float scale, int flip, int N, int H, int W, float* prob_out) {
    DbNet* m = dynamic_cast<DbNet*>(e->model.get());
    if (!m) return set_err(e, DV_ERR_STATE, "handle was not created as a dbnet model");
    if (N <= 0 || H <= 0 || W <= 0 || !prob_out) return set_err(e, DV_ERR_ARG, "dbnet_forward: bad arguments");
    if (m->N != N || m->H != H || m->W != W) {
        // shape change: drop the old plan's buffers and re-plan (static plans => no allocation on the hot path)
        for (void* p : e->owned) cudaFree(p);
        e->owned.clear();
        DV_TRY(build(e, m, N, H, W));
    }
    if (in_nchw) {
        DV_TRY(op_nchw_f32_to_stem(e, in_nchw, N, H, W, m->stem_in.p, m->stem_in.lo));
    } else if (in_u8) {
        DV_TRY(op_u8_to_stem(e, in_u8, N, H, W, mean3, std3, scale, flip, m->stem_in.p, m->stem_in.lo));
    } else {
        return set_err(e, DV_ERR_ARG, "dbnet_forward: no input");
    }
    for (Step& st : m->steps) {
        switch (st.kind) {
            case Step::CONV: DV_TRY(launch_conv(e, st.plan)); break;
            case Step::MAXPOOL: DV_TRY(op_maxpool3x3s2(e, st.a, st.b)); break;
            case Step::DECONV_FINAL: DV_TRY(op_deconv2x2_c1_sigmoid(e, st.a, m->final_w, m->final_w32, m->final_b, prob_out)); break;
        }
    }
    return 0;
}

}  // namespace dv
