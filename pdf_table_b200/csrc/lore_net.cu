// Lore table-structure detector: DLA-34 backbone + DCNv2 up-sampling neck + heads, as a static plan of launches.
// Architecture restated from the reference: DLA / Tree / Root / BasicBlock center_net/modeling_centernet.py:34-402
// (dla34: levels [1,1,1,2,2,1], channels [16,32,64,128,256,512]), DeformConv / IDAUp / DLAUp / DLASeg
// lore/lore_dla_34.py:65-190 (first_level 2, last_level 5, head_conv 256), DCN lore/dcnv2.py:25-86.
// The CPU mirror is oracle/lore_net_ref.py.
//
// What is different from a literal translation
//  * Concatenations never exist as copies: every Root input is one pre-planned NHWC buffer whose channel slices are
//    written directly by the producing kernels (conv epilogue `out_coff`, max-pool `ldo`) and read through strided
//    TMA maps.  The level-3/4 `project` branch whose result the reference computes and then discards (Tree.forward
//    hands it to a sub-Tree that recomputes its own, modeling_centernet.py:266-270) is not evaluated.
//  * DCN = conv_offset_mask (tcgen05 conv, fp32 epilogue) -> k_dcn_im2col (bilinear gather x sigmoid mask, fp16) ->
//    one flat tcgen05 GEMM with K = 9*Cin whose epilogue carries bias + folded BatchNorm + ReLU.
//  * The four small heads (hm, reg, wh, st) run as ONE 3x3 conv 64 -> 1024 and ONE block-diagonal 1x1 -> 24 fp32
//    channels (+ sigmoid on hm): the packed map dv_lore_decode reads.
//  * The two wide heads (ax, cr: 256 channels each, 56 of the network's 338 GFLOP per image, 134 MB of fp32 maps)
//    are never evaluated densely: the reference only ever gathers them at the selected cells' centres / corners, so
//    lore_cell_features() builds the 3x3 patches at those points and runs the two head GEMMs on (5 x cells) rows.
#include <stdlib.h>

#include "engine.h"

namespace dv {

int op_img_to_stem8(Engine* e, const uint8_t* u8, const float* f32, int N, int H, int W, const float* mean3, const float* std3,
                    int flip, __half* out, long long lo, int cpp);
int op_maxpool2x2(Engine* e, const Tensor& in, const Tensor& out);
int op_dcn_im2col(Engine* e, const Tensor& in, const float* om, __half* col, const char* layer);
int op_up_dw_add(Engine* e, const Tensor& in, const float* wt, int f, const Tensor& skip, const Tensor& out, const char* layer);
int op_sigmoid_cols(Engine* e, float* maps, long long rows, int ld, int ncols);
int op_copy_slice(Engine* e, const Tensor& in, const Tensor& out);
int op_cell_offsets(Engine* e, const int32_t* counts, int N, int cap, int32_t* offsets, int32_t* totals, int32_t* overflow);
int op_gather_patch3x3(Engine* e, const Tensor& feat, int K, int cap, const int32_t* counts, const int32_t* offsets,
                       const int32_t* ax_idx, const int32_t* cr_idx, __half* col_ax, __half* col_cr);
int op_logi_combine(Engine* e, const float* ax, const float* cr, int C, int cap, const int32_t* totals, float* out);
int op_gather_pix(Engine* e, const Tensor& fa, const Tensor& fc, int K, int cap, const int32_t* counts, const int32_t* offsets,
                  const int32_t* ax_idx, const int32_t* cr_idx, __half* col_ax, __half* col_cr);

namespace {

constexpr int kCh[6] = {16, 32, 64, 128, 256, 512};
constexpr int kLevels[6] = {1, 1, 1, 2, 2, 1};

struct Step {
    enum Kind { CONV, MAXPOOL, IM2COL, UPADD, SIGMOID, COPY, WINCONV, DCN, MAXPOOL3 } kind;
    ConvPlan plan;
    DcnPlan dcn;
    WinConvPlan win;
    double win_flops = 0;
    Tensor a, b, c;
    const float* wt = nullptr;
    int f = 0;
    std::string name;
};

// fp32x mode (blob entry "precision", weights.pack_lore_dla34(precise=True)): as in dbnet.cu every activation is a split-fp16
// pair and every conv weight a [W_hi | W_lo | W_hi] triple per filter tap; root concatenations are [hi(all) | lo(all)] buffers
// whose channel slices keep the buffer-wide lo offset; the deformable sampling blends hi + lo in fp32 with fp32 weights and
// writes split column rows; the two full-resolution layers run on conv_igemm_tcgen05 (the window-conv kernel keeps fp16
// weights in shared memory).  Head maps land within 1e-3 of the fp32 oracle's range (tests/test_gpu_lore.py).
struct LoreNet : Model {
    Engine* e = nullptr;
    bool precise = false;
    bool plain_up = false;  // CenterNet: DLAUp of plain IDAUp blocks (no DCN), heads hm / v2c / c2v / reg
    int level0_pad = 0;     // 1: the level0 tensor is zero-bordered [N, H + 2, W + 8, 16] with its interior at (+1, +1)
    bool resnet = false;    // Lore wireless: ResNet-18 key-point detector (build_r18); ax / cr end in a 1x1 over their own hidden maps
    Tensor feat_ax, feat_cr;
    int N = 0, H = 0, W = 0;
    std::vector<void*> mem;
    std::vector<Step> steps;
    Tensor stem_in, feat;
    float* om = nullptr;   // [M_max, 32] fp32 conv_offset_mask scratch
    __half* col = nullptr;  // [M_max, 9*C] fp16 deformable-sampling scratch
    size_t col_elems = 0, om_rows = 0;
    float* maps = nullptr;  // [N, H/4, W/4, 24] fp32 (internal copy when the caller passes no buffer)
    std::map<std::string, Tensor> named;
    double flops = 0;
    // sparse ax / cr heads
    int cap = 0, K = 0;
    std::vector<void*> feat_mem;
    std::vector<ConvPlan> feat_plans;
    int32_t *offsets = nullptr, *totals = nullptr, *overflow = nullptr;
    __half *col_ax = nullptr, *col_cr = nullptr, *hid_ax = nullptr, *hid_cr = nullptr;
    float *out_ax = nullptr, *out_cr = nullptr;
    ~LoreNet() override {
        for (void* p : mem) cudaFree(p);
        for (void* p : feat_mem) cudaFree(p);
    }
    int alloc(std::vector<void*>& pool, void** p, size_t bytes, bool zero = false) {
        cudaError_t st = cudaMalloc(p, bytes ? bytes : 16);
        if (st != cudaSuccess) return set_err(e, DV_ERR_CUDA, "lore: cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(st));
        pool.push_back(*p);
        if (zero) cudaMemsetAsync(*p, 0, bytes, e->stream);
        return 0;
    }
    int tensor(Tensor* t, int n, int h, int w, int c, bool zero = false) {
        t->N = n;
        t->H = h;
        t->W = w;
        t->C = c;
        t->ld = 0;
        t->lo = 0;
        void* p = nullptr;
        DV_TRY(alloc(mem, &p, t->elems() * sizeof(__half) * (precise ? 2 : 1), zero));
        t->p = reinterpret_cast<__half*>(p);
        if (precise) {  // [hi(C) | lo(C)] per pixel
            t->ld = 2 * c;
            t->lo = c;
        }
        return 0;
    }
    // the zero-bordered stem image: never a pixel pair -- the lo copy is a second image batch behind the first
    int stem_tensor(Tensor* t, int n, int h, int w, int c) {
        const bool pr = precise;
        precise = false;
        const int rc = tensor(t, pr ? 2 * n : n, h, w, c, /*zero=*/true);
        precise = pr;
        if (rc == 0 && pr) {
            t->N = n;
            t->lo = static_cast<long long>(n) * h * w * c;
        }
        return rc;
    }
    // plan_* allocate their delta tables through the engine: move them into this model's pool
    void adopt(std::vector<void*>& pool) {
        pool.push_back(e->owned.back());
        e->owned.pop_back();
    }
};

int get_conv(Engine* e, const std::string& name, ConvSpec* cs) {
    const BlobTensor* w = e->find(name + ".w");
    const BlobTensor* b = e->find(name + ".b");
    if (!w || !b) return set_err(e, DV_ERR_WEIGHTS, "missing weights for '%s'", name.c_str());
    if (w->dtype != 1 || b->dtype != 0 || w->ndim != 2) return set_err(e, DV_ERR_WEIGHTS, "bad dtype/rank for '%s'", name.c_str());
    cs->w = reinterpret_cast<const __half*>(w->dptr);
    cs->bias = reinterpret_cast<const float*>(b->dptr);
    const int parts = cs->split ? 3 : 1;  // fp32x: [W_hi | W_lo | W_hi] per filter tap
    const int taps = cs->KH * cs->KW * parts;
    if (cs->stem) {
        if (static_cast<int>(w->dims[0]) != cs->Cout || static_cast<int>(w->dims[1]) != 448 * parts)
            return set_err(e, DV_ERR_WEIGHTS, "'%s': stride-1 stem weight must be [%d,%d]", name.c_str(), cs->Cout, 448 * parts);
        cs->Cin_pad = 64;
        cs->BK = 64;
    } else {
        if (static_cast<int>(w->dims[0]) != cs->Cout || (w->dims[1] % taps) != 0)
            return set_err(e, DV_ERR_WEIGHTS, "'%s': weight shape [%u,%u] does not match Cout=%d taps=%d", name.c_str(), w->dims[0],
                           w->dims[1], cs->Cout, taps);
        cs->Cin_pad = static_cast<int>(w->dims[1]) / taps;
        if (cs->split && cs->KH == 1 && cs->stride == 1) cs->Cin_pad *= 3;  // flat GEMM: plan_linear counts all three parts
        cs->BK = (cs->Cin_pad % 64 == 0) ? 64 : (cs->Cin_pad % 32 == 0) ? 32 : 16;
    }
    if (b->dims[0] < static_cast<uint32_t>((cs->Cout + 255) / 256 * 256))
        return set_err(e, DV_ERR_WEIGHTS, "'%s': bias not padded to 256", name.c_str());
    return 0;
}

EpiSpec epi(const Tensor& out, int act, const Tensor* res = nullptr) {
    EpiSpec es;
    es.out = out.p;
    es.out_ld = out.ldc();
    es.split_off = static_cast<int>(out.lo);
    es.act = act;
    if (res) {
        es.res = res->p;
        es.res_mode = RES_SAME;
        es.res_ld = res->ldc();
        es.res_lo = static_cast<int>(res->lo);
    }
    return es;
}

int add_conv(LoreNet* m, const std::string& name, const Tensor& in, int cout, int k, int stride, const EpiSpec& es, bool stem = false) {
    ConvSpec cs;
    cs.KH = cs.KW = k;
    cs.stride = stride;
    cs.pad = k / 2;
    cs.Cin = stem ? 3 : in.C;
    cs.Cout = cout;
    cs.stem = stem;
    cs.split = m->precise;
    DV_TRY(get_conv(m->e, name, &cs));
    const int Ho = stem ? in.H - 6 : in.H / stride, Wo = stem ? in.W - 8 : in.W / stride;
    Step st;
    st.kind = Step::CONV;
    st.name = name;
    DV_TRY(plan_conv(m->e, in, cs, es, Ho, Wo, &st.plan, st.name.c_str()));
    m->adopt(m->mem);
    m->flops += st.plan.flops;
    m->steps.push_back(st);
    return 0;
}

int add_pool(LoreNet* m, const Tensor& in, const Tensor& out) {
    Step st;
    st.kind = Step::MAXPOOL;
    st.a = in;
    st.b = out;
    m->steps.push_back(st);
    return 0;
}

// BasicBlock (modeling_centernet.py:34-71): relu(bn2(conv2(relu(bn1(conv1(x))))) + residual) -> dst
int add_block(LoreNet* m, const std::string& p, const Tensor& x, int cout, int stride, const Tensor& residual, const Tensor& dst) {
    Tensor t;
    DV_TRY(m->tensor(&t, x.N, x.H / stride, x.W / stride, cout));
    DV_TRY(add_conv(m, p + ".conv1", x, cout, 3, stride, epi(t, ACT_RELU)));
    DV_TRY(add_conv(m, p + ".conv2", t, cout, 3, 1, epi(dst, ACT_RELU, &residual)));
    return 0;
}

// A levels == 1 Tree (modeling_centernet.py:186-287).  `root` is the pre-planned Root input [x2 | x1 | children...];
// `bottom` = max-pooled input (already planned by the caller when stride > 1), `dst` receives the Root output.
int add_tree1(LoreNet* m, const std::string& p, const Tensor& x, int cin, int cout, int stride, const Tensor& bottom, const Tensor& root,
              const Tensor& dst) {
    Tensor residual = bottom;
    if (cin != cout) {
        DV_TRY(m->tensor(&residual, bottom.N, bottom.H, bottom.W, cout));
        DV_TRY(add_conv(m, p + ".project", bottom, cout, 1, 1, epi(residual, ACT_NONE)));
    }
    const Tensor x1 = root.slice(cout, cout), x2 = root.slice(0, cout);
    DV_TRY(add_block(m, p + ".tree1", x, cout, stride, residual, x1));
    DV_TRY(add_block(m, p + ".tree2", x1, cout, 1, x1, x2));
    DV_TRY(add_conv(m, p + ".root", root, cout, 1, 1, epi(dst, ACT_RELU)));
    return 0;
}

// One DLA level (a Tree with stride 2) -> dense output
int add_level(LoreNet* m, int lvl, const Tensor& x, Tensor* out) {
    const std::string p = "level" + std::to_string(lvl);
    const int cin = kCh[lvl - 1], cout = kCh[lvl], levels = kLevels[lvl];
    const bool level_root = lvl > 2;
    const int Ho = x.H / 2, Wo = x.W / 2;
    DV_TRY(m->tensor(out, x.N, Ho, Wo, cout));
    if (levels == 1) {
        Tensor root;
        DV_TRY(m->tensor(&root, x.N, Ho, Wo, 2 * cout + (level_root ? cin : 0)));
        Tensor bottom;
        if (level_root) bottom = root.slice(2 * cout, cin);
        else DV_TRY(m->tensor(&bottom, x.N, Ho, Wo, cin));
        DV_TRY(add_pool(m, x, bottom));
        DV_TRY(add_tree1(m, p, x, cin, cout, 2, bottom, root, *out));
    } else {
        // levels == 2: tree1 = Tree(1, cin -> cout, stride 2), tree2 = Tree(1, cout -> cout) whose Root also takes
        // [bottom, x1]; root2 = [x2b | x1b | bottom | x1]
        Tensor root2, root1;
        DV_TRY(m->tensor(&root2, x.N, Ho, Wo, 2 * cout + cin + cout));
        DV_TRY(m->tensor(&root1, x.N, Ho, Wo, 2 * cout));
        const Tensor bottom = root2.slice(2 * cout, cin), x1 = root2.slice(2 * cout + cin, cout);
        DV_TRY(add_pool(m, x, bottom));
        DV_TRY(add_tree1(m, p + ".tree1", x, cin, cout, 2, bottom, root1, x1));
        DV_TRY(add_tree1(m, p + ".tree2", x1, cout, cout, 1, x1, root2, *out));
    }
    m->named[p] = *out;
    return 0;
}

// DeformConv (lore_dla_34.py:65-85): DCN + BN + ReLU -> dst (dense)
int add_dcn(LoreNet* m, const std::string& p, const Tensor& x, int cout, Tensor* dst) {
    DV_TRY(m->tensor(dst, x.N, x.H, x.W, cout));
    const size_t rows = static_cast<size_t>(x.N) * x.H * x.W;
    if (rows > m->om_rows || rows * 9 * x.C * (m->precise ? 2 : 1) > m->col_elems)
        return set_err(m->e, DV_ERR_STATE, "lore: DCN scratch too small for %s", p.c_str());
    {
        EpiSpec es;
        es.out = m->om;
        es.out_ld = 32;
        es.out_f32 = 1;
        DV_TRY(add_conv(m, p + ".om", x, 32, 3, 1, es));
    }
    if (!m->precise && dcn_fused_enabled() && x.C % 64 == 0 && (cout == 64 || cout == 128 || cout == 256)) {
        // one kernel: sampling producers -> swizzled shared-memory A tiles -> tcgen05 GEMM -> bias + ReLU (dcn_fused.cuh)
        ConvSpec cs;
        cs.KH = cs.KW = 1;
        cs.Cout = cout;
        DV_TRY(get_conv(m->e, p + ".dcn", &cs));
        if (cs.Cin_pad != 9 * x.C) return set_err(m->e, DV_ERR_WEIGHTS, "'%s.dcn': K %d != 9*%d", p.c_str(), cs.Cin_pad, x.C);
        Step st;
        st.kind = Step::DCN;
        st.name = p + ".dcn";
        DV_TRY(plan_dcn(m->e, x, m->om, cs.w, cs.bias, cout, ACT_RELU, *dst, &st.dcn, st.name.c_str()));
        m->flops += st.dcn.flops;
        m->steps.push_back(st);
        return 0;
    }
    {
        Step st;
        st.kind = Step::IM2COL;
        st.a = x;
        st.name = p + ".sample";
        m->steps.push_back(st);
        m->flops += 0;
    }
    {
        ConvSpec cs;
        cs.KH = cs.KW = 1;
        cs.Cout = cout;
        cs.split = m->precise;
        DV_TRY(get_conv(m->e, p + ".dcn", &cs));
        if (cs.Cin_pad != 9 * x.C * (m->precise ? 3 : 1)) return set_err(m->e, DV_ERR_WEIGHTS, "'%s.dcn': K %d != 9*%d", p.c_str(), cs.Cin_pad, x.C);
        cs.Cin = 9 * x.C;
        cs.BK = 64;
        cs.flat = true;
        Step st;
        st.kind = Step::CONV;
        st.name = p + ".dcn";
        DV_TRY(plan_linear(m->e, m->col, static_cast<int>(rows), 9 * x.C, cs, epi(*dst, ACT_RELU), &st.plan, st.name.c_str()));
        m->adopt(m->mem);
        m->flops += st.plan.flops;
        m->steps.push_back(st);
    }
    return 0;
}

// IDAUp.forward (lore_dla_34.py:104-110) over layers[startp+1 .. endp)
int add_ida(LoreNet* m, const std::string& p, std::vector<Tensor>& layers, int startp, int endp, int o, const int* up_f) {
    for (int i = startp + 1; i < endp; ++i) {
        const int j = i - startp, f = up_f[j];
        Tensor proj, sum, node;
        DV_TRY(add_dcn(m, p + ".proj_" + std::to_string(j), layers[i], o, &proj));
        const BlobTensor* w = m->e->find(p + ".up_" + std::to_string(j) + ".w");
        if (!w || w->dtype != 0 || w->nbytes != static_cast<uint64_t>(4 * f * f * o) * 4)
            return set_err(m->e, DV_ERR_WEIGHTS, "missing / bad '%s.up_%d.w'", p.c_str(), j);
        DV_TRY(m->tensor(&sum, proj.N, proj.H * f, proj.W * f, o));
        Step st;
        st.kind = Step::UPADD;
        st.a = proj;
        st.b = layers[i - 1];
        st.c = sum;
        st.wt = reinterpret_cast<const float*>(w->dptr);
        st.f = f;
        st.name = p + ".up_" + std::to_string(j);
        m->steps.push_back(st);
        DV_TRY(add_dcn(m, p + ".node_" + std::to_string(j), sum, o, &node));
        layers[i] = node;
    }
    return 0;
}

// CenterNet IDAUp.forward (center_net/modeling_centernet.py:555-570): layers[k] = up_k(proj_k(layers[k])), then
// x = node_k(cat[x, layers[k]]).  cat buffers: [x | up(proj(l_k))]; x of the first node is a copy of layers[0] (it is also
// read by later IDAs and, for stride-2 convs, must stay dense), later x are written straight into the next cat by the node conv.
int add_plain_ida(LoreNet* m, const std::string& p, std::vector<Tensor>& ls, int o, const int* up_f, std::vector<Tensor>* y) {
    Engine* e = m->e;
    const int n = static_cast<int>(ls.size());
    const Tensor& l0 = ls[0];
    std::vector<Tensor> cat(n);
    for (int k = 1; k < n; ++k) DV_TRY(m->tensor(&cat[k], l0.N, l0.H, l0.W, 2 * o));
    {
        Step st;
        st.kind = Step::COPY;
        st.a = l0;
        st.b = cat[1].slice(0, o);
        m->steps.push_back(st);
    }
    for (int k = 1; k < n; ++k) {
        Tensor t = ls[k];
        const std::string pk = p + ".proj_" + std::to_string(k);
        if (e->find(pk + ".w")) {
            Tensor pr;
            DV_TRY(m->tensor(&pr, t.N, t.H, t.W, o));
            DV_TRY(add_conv(m, pk, t, o, 1, 1, epi(pr, ACT_RELU)));
            t = pr;
        } else if (t.C != o) {
            return set_err(e, DV_ERR_WEIGHTS, "missing '%s'", pk.c_str());
        }
        const int f = up_f[k];
        const BlobTensor* w = e->find(p + ".up_" + std::to_string(k) + ".w");
        if (f == 1 || !w || w->dtype != 0 || w->nbytes != static_cast<uint64_t>(4 * f * f * o) * 4)
            return set_err(e, DV_ERR_WEIGHTS, "missing / bad '%s.up_%d.w'", p.c_str(), k);
        Step st;
        st.kind = Step::UPADD;
        st.a = t;
        st.b = Tensor();  // no skip: plain up-sampling into the concatenation slice
        st.c = cat[k].slice(o, o);
        st.wt = reinterpret_cast<const float*>(w->dptr);
        st.f = f;
        st.name = p + ".up_" + std::to_string(k);
        m->steps.push_back(st);
    }
    for (int k = 1; k < n; ++k) {
        Tensor out;
        DV_TRY(m->tensor(&out, l0.N, l0.H, l0.W, o));
        DV_TRY(add_conv(m, p + ".node_" + std::to_string(k), cat[k], o, 3, 1, epi(out, ACT_RELU)));
        if (k + 1 < n) {
            Step st;
            st.kind = Step::COPY;
            st.a = out;
            st.b = cat[k + 1].slice(0, o);
            m->steps.push_back(st);
        }
        y->push_back(out);
    }
    return 0;
}

int build(Engine* e, LoreNet* m, int N, int H, int W) {
    if ((H % 32) || (W % 32)) return set_err(e, DV_ERR_ARG, "lore: H and W must be multiples of 32 (got %dx%d)", H, W);
    for (void* p : m->mem) cudaFree(p);
    m->mem.clear();
    m->steps.clear();
    m->named.clear();
    m->flops = 0;
    m->e = e;
    m->N = N;
    m->H = H;
    m->W = W;
    DV_TRY(m->stem_tensor(&m->stem_in, N, H + 6, W + 8, 8));
    // DCN scratch: the largest sampling matrix is 9*64 channels at stride 4 (or 9*128 at stride 8: same size)
    m->om_rows = static_cast<size_t>(N) * (H / 4) * (W / 4);
    m->col_elems = m->plain_up ? 0 : m->om_rows * 9 * 64 * (m->precise ? 2 : 1);
    {
        void* p = nullptr;
        if (!m->plain_up) {
            DV_TRY(m->alloc(m->mem, &p, m->om_rows * 32 * sizeof(float)));
            m->om = reinterpret_cast<float*>(p);
            DV_TRY(m->alloc(m->mem, &p, m->col_elems * sizeof(__half)));
            m->col = reinterpret_cast<__half*>(p);
        }
        DV_TRY(m->alloc(m->mem, &p, m->om_rows * 24 * sizeof(float)));
        m->maps = reinterpret_cast<float*>(p);
    }
    Tensor b0, l0, l1;
    DV_TRY(m->tensor(&l1, N, H / 2, W / 2, 32));
    static const bool win_env = !(getenv("DV_WINCONV") && atoi(getenv("DV_WINCONV")) == 0);
    static const bool win_l1_env = !(getenv("DV_WINCONV_L1") && atoi(getenv("DV_WINCONV_L1")) == 0);
    const bool use_win = win_env && !m->precise;
    // level1 (3x3 stride 2, 16 -> 32) also runs on conv_win_tcgen05 when its window-packed filter is in the blob: level0 then writes
    // into a zero-bordered buffer (interior at +1,+1) -- through the TMA patch mode its 32-byte rows made it TMA-row-rate bound
    const bool win_l1 = use_win && win_l1_env && e->find("level1.win.w") != nullptr;
    m->level0_pad = win_l1 ? 1 : 0;
    if (win_l1) DV_TRY(m->tensor(&l0, N, H + 2, W + 8, 16, /*zero=*/true));
    else DV_TRY(m->tensor(&l0, N, H, W, 16));
    if (use_win) {
        // the two full-resolution 16-channel layers run on conv_win_tcgen05 (load/store producer, resident filter):
        // base writes into a zero-bordered buffer (interior at +1,+1) so that level0 reads 4-pixel x 16-channel windows
        DV_TRY(m->tensor(&b0, N, H + 2, W + 8, 16, /*zero=*/true));
        const char* names[2] = {"base", win_patch_enabled() ? "level0.winp" : "level0.win"};
        for (int i = 0; i < 2; ++i) {
            const BlobTensor* w = e->find(std::string(names[i]) + ".w");
            const BlobTensor* b = e->find(std::string(names[i]) + ".b");
            const uint32_t kcols = i == 0 ? 448 : 192;
            if (!w || !b || w->dtype != 1 || b->dtype != 0 || w->dims[0] != 16 || w->dims[1] != kcols)
                return set_err(e, DV_ERR_WEIGHTS, "missing / bad '%s' (re-pack the weights with this version)", names[i]);
            Step st;
            st.kind = Step::WINCONV;
            st.name = i == 0 ? "base" : "level0";
            DV_TRY(plan_win_conv(e, i == 0 ? m->stem_in : b0, 1, i == 0 ? 7 : 3, H, W, reinterpret_cast<const __half*>(w->dptr),
                                 reinterpret_cast<const float*>(b->dptr), 16, ACT_RELU, i == 0 ? b0 : l0, (i == 0 || win_l1) ? 1 : 0, &st.win, st.name.c_str()));
            st.win_flops = 2.0 * N * H * W * (i == 0 ? 147.0 : 144.0) * 16;
            m->flops += st.win_flops;
            m->steps.push_back(st);
        }
    } else {
        DV_TRY(m->tensor(&b0, N, H, W, 16));
        DV_TRY(add_conv(m, "base", m->stem_in, 16, 7, 1, epi(b0, ACT_RELU), /*stem=*/true));
        DV_TRY(add_conv(m, "level0", b0, 16, 3, 1, epi(l0, ACT_RELU)));
    }
    if (win_l1) {
        const BlobTensor* w = e->find("level1.win.w");
        const BlobTensor* b = e->find("level1.win.b");
        if (!b || w->dtype != 1 || b->dtype != 0 || w->dims[0] != 32 || w->dims[1] != 192) return set_err(e, DV_ERR_WEIGHTS, "bad 'level1.win'");
        Step st;
        st.kind = Step::WINCONV;
        st.name = "level1";
        DV_TRY(plan_win_conv(e, l0, 2, 3, H / 2, W / 2, reinterpret_cast<const __half*>(w->dptr), reinterpret_cast<const float*>(b->dptr), 32, ACT_RELU,
                             l1, 0, &st.win, st.name.c_str()));
        st.win_flops = 2.0 * N * (H / 2) * (W / 2) * 144.0 * 32;
        m->flops += st.win_flops;
        m->steps.push_back(st);
    } else {
        DV_TRY(add_conv(m, "level1", l0, 32, 3, 2, epi(l1, ACT_RELU)));
    }
    m->named["level0"] = l0;
    m->named["level1"] = l1;
    std::vector<Tensor> layers(4);
    Tensor x = l1;
    for (int lvl = 2; lvl < 6; ++lvl) {
        DV_TRY(add_level(m, lvl, x, &layers[lvl - 2]));
        x = layers[lvl - 2];
    }
    if (m->plain_up) {
        // CenterNet DLAUp.forward (center_net/modeling_centernet.py:589-597): ida_i over layers[-i-2:], layers[-i-1:] = y
        const int up2[4] = {1, 2, 2, 2};
        const int outs[3] = {256, 128, 64};
        for (int i = 0; i < 3; ++i) {
            std::vector<Tensor> ls(layers.end() - i - 2, layers.end());
            std::vector<Tensor> y;
            DV_TRY(add_plain_ida(m, "dla_up.ida_" + std::to_string(i), ls, outs[i], up2, &y));
            for (size_t k = 0; k < y.size(); ++k) layers[layers.size() - y.size() + k] = y[k];
            m->feat = y.back();
        }
    }
    // DLAUp.forward (lore_dla_34.py:130-137): channels [64,128,256,512], scales [1,2,4,8]
    std::vector<Tensor> out(1, layers[3]);
    if (!m->plain_up) {
        const int up2[4] = {1, 2, 2, 2};
        DV_TRY(add_ida(m, "dla_up.ida_0", layers, 2, 4, 256, up2));
        out.insert(out.begin(), layers[3]);
        DV_TRY(add_ida(m, "dla_up.ida_1", layers, 1, 4, 128, up2));
        out.insert(out.begin(), layers[3]);
        DV_TRY(add_ida(m, "dla_up.ida_2", layers, 0, 4, 64, up2));
        out.insert(out.begin(), layers[3]);
    }
    // DLASeg.forward (:176-189): y = out[0:3]; ida_up(y, 0, 3) with up factors [1,2,4]
    if (!m->plain_up) {
        std::vector<Tensor> y(out.begin(), out.begin() + 3);
        const int upf[3] = {1, 2, 4};
        DV_TRY(add_ida(m, "ida_up", y, 0, 3, 64, upf));
        m->feat = y[2];
    }
    m->named["feat"] = m->feat;
    // small heads
    Tensor hid;
    DV_TRY(m->tensor(&hid, N, m->feat.H, m->feat.W, 1024));
    DV_TRY(add_conv(m, "heads.conv", m->feat, 1024, 3, 1, epi(hid, ACT_RELU)));
    {
        EpiSpec es;
        es.out = m->maps;  // re-pointed per call when the caller supplies its own buffer
        es.out_ld = 24;
        es.out_f32 = 1;
        DV_TRY(add_conv(m, "heads.out", hid, 24, 1, 1, es));
    }
    {
        Step st;
        st.kind = Step::SIGMOID;
        st.name = "hm.sigmoid";
        m->steps.push_back(st);
    }
    return 0;
}

// ------------------------------------------------------------------------------------------------ Lore wireless (ResNet-18)
// LoreDetectModel (lore/lore_detector.py:148-389).  The 7x7 stride-2 stem reads the zero-bordered 4-channel image through the
// overlapping-window view (A_STEM, as DBNet's); every stage is entered with stride 2 (:180-187); each ConvTranspose 4x4 s2 + BN +
// ReLU is a 3x3 conv to 4 x 256 channels with the pixel-shuffle store (weights.deconv4x4_as_conv3x3) and the 1x1 `adaption`
// lateral takes it as its epilogue residual; the six heads' first convs are one 3x3 conv to 6 x 64 channels, their 64 -> 64 convs
// ping-pong between two 384-wide buffers on channel slices, the small heads end in one block-diagonal 1x1 to the packed 24-wide
// fp32 map, and `ax` / `cr` leave their 64-channel hidden maps for the sparse 1x1 at the decoded points (build_feat).
int add_stem_s2(LoreNet* m, const Tensor& in, const Tensor& out) {
    Engine* e = m->e;
    const BlobTensor* w = e->find("stem.w");
    const BlobTensor* b = e->find("stem.b");
    const int parts = m->precise ? 3 : 1;
    if (!w || !b || w->dtype != 1 || b->dtype != 0 || w->ndim != 2 || w->dims[0] != 64 || static_cast<int>(w->dims[1]) != 224 * parts)
        return set_err(e, DV_ERR_WEIGHTS, "lore_resnet18: stem weight must be [64,%d]", 224 * parts);
    ConvSpec cs;
    cs.KH = cs.KW = 7;
    cs.stride = 2;
    cs.pad = 3;
    cs.Cin = 3;
    cs.Cout = 64;
    cs.stem = true;
    cs.split = m->precise;
    cs.BK = 32;
    cs.Cin_pad = 32;
    cs.w = reinterpret_cast<const __half*>(w->dptr);
    cs.bias = reinterpret_cast<const float*>(b->dptr);
    Step st;
    st.kind = Step::CONV;
    st.name = "stem";
    DV_TRY(plan_conv(e, in, cs, epi(out, ACT_RELU), out.H, out.W, &st.plan, "stem"));
    m->adopt(m->mem);
    m->flops += st.plan.flops;
    m->steps.push_back(st);
    return 0;
}

int build_r18(Engine* e, LoreNet* m, int N, int H, int W) {
    if ((H % 64) || (W % 64)) return set_err(e, DV_ERR_ARG, "lore_resnet18: H and W must be multiples of 64 (got %dx%d)", H, W);
    for (void* p : m->mem) cudaFree(p);
    m->mem.clear();
    m->steps.clear();
    m->named.clear();
    m->flops = 0;
    m->e = e;
    m->N = N;
    m->H = H;
    m->W = W;
    DV_TRY(m->stem_tensor(&m->stem_in, N, H + 6, W + 8, 4));
    m->om_rows = static_cast<size_t>(N) * (H / 4) * (W / 4);
    {
        void* p = nullptr;
        DV_TRY(m->alloc(m->mem, &p, m->om_rows * 24 * sizeof(float)));
        m->maps = reinterpret_cast<float*>(p);
    }
    Tensor c1, x[5];
    DV_TRY(m->tensor(&c1, N, H / 2, W / 2, 64));
    DV_TRY(m->tensor(&x[0], N, H / 4, W / 4, 64));
    DV_TRY(add_stem_s2(m, m->stem_in, c1));
    {
        Step st;
        st.kind = Step::MAXPOOL3;
        st.a = c1;
        st.b = x[0];
        m->steps.push_back(st);
    }
    m->named["c1"] = c1;
    m->named["x0"] = x[0];
    const int planes[4] = {64, 128, 256, 256};
    Tensor y = x[0];
    for (int L = 0; L < 4; ++L) {
        for (int B = 0; B < 2; ++B) {
            const int stride = B == 0 ? 2 : 1;
            const std::string pre = "layer" + std::to_string(L + 1) + "." + std::to_string(B);
            Tensor t, o, ds;
            DV_TRY(m->tensor(&t, N, y.H / stride, y.W / stride, planes[L]));
            DV_TRY(m->tensor(&o, N, y.H / stride, y.W / stride, planes[L]));
            DV_TRY(add_conv(m, pre + ".conv1", y, planes[L], 3, stride, epi(t, ACT_RELU)));
            const Tensor* res = &y;
            if (B == 0) {
                DV_TRY(m->tensor(&ds, N, y.H / stride, y.W / stride, planes[L]));
                DV_TRY(add_conv(m, pre + ".down", y, planes[L], 1, stride, epi(ds, ACT_NONE)));
                res = &ds;
            }
            DV_TRY(add_conv(m, pre + ".conv2", t, planes[L], 3, 1, epi(o, ACT_RELU, res)));
            y = o;
        }
        x[L + 1] = y;
        m->named["x" + std::to_string(L + 1)] = y;
    }
    // top-down: up = ConvTranspose 4x4 s2 + BN + ReLU (3x3 conv + pixel shuffle), lateral = adaption 1x1 + up as residual
    Tensor top = x[4];
    const char* lateral[4] = {"adaption3", "adaption2", "adaption1", "adaption0"};
    for (int i = 0; i < 4; ++i) {
        const Tensor& skip = x[3 - i];
        Tensor up, sum;
        DV_TRY(m->tensor(&up, N, skip.H, skip.W, 256));
        DV_TRY(m->tensor(&sum, N, skip.H, skip.W, 256));
        EpiSpec es = epi(up, ACT_RELU);
        es.out_mode = OUT_SHUF2;
        DV_TRY(add_conv(m, "up" + std::to_string(i + 1), top, 1024, 3, 1, es));
        {  // algorithmic work of the transposed conv: 4 of the 9 taps per output parity; the zero taps are not counted
            ConvPlan& pl = m->steps.back().plan;
            m->flops -= pl.flops * (5.0 / 9.0);
            pl.flops *= 4.0 / 9.0;
        }
        DV_TRY(add_conv(m, lateral[i], skip, 256, 1, 1, epi(sum, ACT_NONE, &up)));
        m->named["x" + std::to_string(3 - i) + "_"] = sum;
        top = sum;
    }
    DV_TRY(m->tensor(&m->feat, N, H / 4, W / 4, 256));
    DV_TRY(add_conv(m, "adaptionU1", top, 256, 1, 1, epi(m->feat, ACT_NONE)));
    m->named["feat"] = m->feat;
    // heads: channel slices [hm | reg | wh | st | ax | cr] of two 384-wide buffers
    Tensor A, B;
    DV_TRY(m->tensor(&A, N, H / 4, W / 4, 384));
    DV_TRY(m->tensor(&B, N, H / 4, W / 4, 384));
    DV_TRY(add_conv(m, "heads.conv1", m->feat, 384, 3, 1, epi(A, ACT_RELU)));
    const char* hn[6] = {"hm", "reg", "wh", "st", "ax", "cr"};
    for (int h = 0; h < 6; ++h) {
        if (h == 1) continue;  // reg: conv3x3 + 1x1 only
        const Tensor a = A.slice(64 * h, 64), b = B.slice(64 * h, 64);
        const std::string pre = std::string("heads.") + hn[h];
        DV_TRY(add_conv(m, pre + ".2", a, 64, 3, 1, epi(b, ACT_RELU)));
        DV_TRY(add_conv(m, pre + ".4", b, 64, 3, 1, epi(a, ACT_RELU)));
        DV_TRY(add_conv(m, pre + ".6", a, 64, 3, 1, epi(b, ACT_RELU)));
    }
    {
        Step st;
        st.kind = Step::COPY;
        st.name = "reg.hidden";
        st.a = A.slice(64, 64);
        st.b = B.slice(64, 64);
        m->steps.push_back(st);
    }
    {
        EpiSpec es;
        es.out = m->maps;  // re-pointed per call when the caller supplies its own buffer
        es.out_ld = 24;
        es.out_f32 = 1;
        DV_TRY(add_conv(m, "heads.out", B.slice(0, 256), 24, 1, 1, es));
    }
    {
        Step st;
        st.kind = Step::SIGMOID;
        st.name = "hm.sigmoid";
        m->steps.push_back(st);
    }
    m->feat_ax = B.slice(256, 64);
    m->feat_cr = B.slice(320, 64);
    return 0;
}

int get_flat(Engine* e, const std::string& name, int K, int N, ConvSpec* cs) {
    cs->KH = cs->KW = 1;
    cs->Cout = N;
    DV_TRY(get_conv(e, name, cs));
    if (cs->Cin_pad != K * (cs->split ? 3 : 1)) return set_err(e, DV_ERR_WEIGHTS, "'%s': K %d != %d", name.c_str(), cs->Cin_pad, K);
    cs->Cin = K;
    cs->flat = true;
    return 0;
}

int build_feat(Engine* e, LoreNet* m, int K, int cap) {
    for (void* p : m->feat_mem) cudaFree(p);
    m->feat_mem.clear();
    m->feat_plans.clear();
    m->K = K;
    m->cap = cap;
    const int C = 64, D = 256;
    void* p = nullptr;
    DV_TRY(m->alloc(m->feat_mem, &p, (m->N + 1 + 2 + 1) * sizeof(int32_t), true));
    m->offsets = reinterpret_cast<int32_t*>(p);
    m->totals = m->offsets + m->N + 1;
    m->overflow = m->totals + 2;
    const size_t sp = m->precise ? 2 : 1;  // fp32x: patch rows and hidden rows are [hi | lo] pairs
    const int taps = m->resnet ? 1 : 9;   // ResNet-18 detector: the heads end in a 1x1 over their own hidden pixel
    DV_TRY(m->alloc(m->feat_mem, &p, static_cast<size_t>(cap) * taps * C * 2 * sp, true));
    m->col_ax = reinterpret_cast<__half*>(p);
    DV_TRY(m->alloc(m->feat_mem, &p, static_cast<size_t>(cap) * 4 * taps * C * 2 * sp, true));
    m->col_cr = reinterpret_cast<__half*>(p);
    DV_TRY(m->alloc(m->feat_mem, &p, static_cast<size_t>(cap) * D * 2 * sp, true));
    m->hid_ax = reinterpret_cast<__half*>(p);
    DV_TRY(m->alloc(m->feat_mem, &p, static_cast<size_t>(cap) * 4 * D * 2 * sp, true));
    m->hid_cr = reinterpret_cast<__half*>(p);
    DV_TRY(m->alloc(m->feat_mem, &p, static_cast<size_t>(cap) * D * 4, true));
    m->out_ax = reinterpret_cast<float*>(p);
    DV_TRY(m->alloc(m->feat_mem, &p, static_cast<size_t>(cap) * 4 * D * 4, true));
    m->out_cr = reinterpret_cast<float*>(p);
    const char* heads[2] = {"ax", "cr"};
    for (int h = 0; h < 2; ++h) {
        const int rows = h == 0 ? cap : 4 * cap;
        const int* m_dyn = m->totals + h;
        if (m->resnet) {
            ConvSpec c;
            c.split = m->precise;
            DV_TRY(get_flat(e, std::string(heads[h]) + ".out", C, D, &c));
            EpiSpec es;
            es.out = h == 0 ? m->out_ax : m->out_cr;
            es.out_ld = D;
            es.out_f32 = 1;
            es.m_dyn = m_dyn;
            ConvPlan pl;
            DV_TRY(plan_linear(e, h == 0 ? m->col_ax : m->col_cr, rows, C, c, es, &pl, (std::string(heads[h]) + ".out").c_str()));
            m->adopt(m->feat_mem);
            m->feat_plans.push_back(pl);
            continue;
        }
        ConvSpec c1, c2;
        c1.split = c2.split = m->precise;
        DV_TRY(get_flat(e, std::string(heads[h]) + ".conv", 9 * C, D, &c1));
        DV_TRY(get_flat(e, std::string(heads[h]) + ".out", D, D, &c2));
        EpiSpec e1, e2;
        e1.out = h == 0 ? m->hid_ax : m->hid_cr;
        e1.out_ld = D * static_cast<int>(sp);
        e1.split_off = m->precise ? D : 0;
        e1.act = ACT_RELU;
        e1.m_dyn = m_dyn;
        e2.out = h == 0 ? m->out_ax : m->out_cr;
        e2.out_ld = D;
        e2.out_f32 = 1;
        e2.m_dyn = m_dyn;
        ConvPlan p1, p2;
        DV_TRY(plan_linear(e, h == 0 ? m->col_ax : m->col_cr, rows, 9 * C, c1, e1, &p1, (std::string(heads[h]) + ".conv").c_str()));
        m->adopt(m->feat_mem);
        DV_TRY(plan_linear(e, h == 0 ? m->hid_ax : m->hid_cr, rows, D, c2, e2, &p2, (std::string(heads[h]) + ".out").c_str()));
        m->adopt(m->feat_mem);
        m->feat_plans.push_back(p1);
        m->feat_plans.push_back(p2);
    }
    return 0;
}

}  // namespace

int lore_create(Engine* e) {
    LoreNet* m = new LoreNet();
    m->e = e;
    m->plain_up = e->kind == "centernet_dla34";
    m->resnet = e->kind == "lore_resnet18";
    m->precise = e->find("precision") != nullptr;
    e->model.reset(m);
    return 0;
}

double lore_flops(Engine* e) {
    LoreNet* m = dynamic_cast<LoreNet*>(e->model.get());
    return m ? m->flops : 0.0;
}

int lore_debug_tensor(Engine* e, const char* name, float* out_nchw, int* dims4) {
    LoreNet* m = dynamic_cast<LoreNet*>(e->model.get());
    if (!m) return set_err(e, DV_ERR_STATE, "not a lore handle");
    auto it = m->named.find(name);
    if (it == m->named.end()) return set_err(e, DV_ERR_ARG, "no tensor named '%s'", name);
    const Tensor& t = it->second;
    if (t.ld != 0 && t.lo == 0) return set_err(e, DV_ERR_UNSUPPORTED, "tensor '%s' is a slice", name);
    const bool pad = m->level0_pad && std::string(name) == "level0";
    if (dims4) {
        dims4[0] = t.N;
        dims4[1] = t.C;
        dims4[2] = pad ? t.H - 2 : t.H;
        dims4[3] = pad ? t.W - 8 : t.W;
    }
    if (out_nchw) {
        if (pad) return op_nhwc_f16_to_nchw_f32(e, t.p, t.N, t.C, t.H - 2, t.W - 8, out_nchw, t.ldc(), static_cast<int>(t.lo), t.H, t.W, 1, 1);
        return op_nhwc_f16_to_nchw_f32(e, t.p, t.N, t.C, t.H, t.W, out_nchw, t.ldc(), static_cast<int>(t.lo));
    }
    return 0;
}

int lore_detect_forward(Engine* e, const float* in_nchw, const uint8_t* in_u8, const float* mean3, const float* std3, int flip, int N,
                        int H, int W, float* maps_out) {
    LoreNet* m = dynamic_cast<LoreNet*>(e->model.get());
    if (!m) return set_err(e, DV_ERR_STATE, "handle was not created as a lore_dla34 / centernet_dla34 model");
    if (N <= 0 || H <= 0 || W <= 0) return set_err(e, DV_ERR_ARG, "lore_detect_forward: bad arguments");
    if (m->N != N || m->H != H || m->W != W) DV_TRY(m->resnet ? build_r18(e, m, N, H, W) : build(e, m, N, H, W));
    if (!in_nchw && !in_u8) return set_err(e, DV_ERR_ARG, "lore_detect_forward: no input");
    DV_TRY(op_img_to_stem8(e, in_u8, in_nchw, N, H, W, mean3, std3, flip, m->stem_in.p, m->stem_in.lo, m->resnet ? 4 : 8));
    float* maps = maps_out ? maps_out : m->maps;
    for (Step& st : m->steps) {
        switch (st.kind) {
            case Step::CONV:
                if (st.name == "heads.out") st.plan.prm.out = maps;
                DV_TRY(launch_conv(e, st.plan));
                break;
            case Step::MAXPOOL: DV_TRY(op_maxpool2x2(e, st.a, st.b)); break;
            case Step::MAXPOOL3: DV_TRY(op_maxpool3x3s2(e, st.a, st.b)); break;
            case Step::IM2COL: DV_TRY(op_dcn_im2col(e, st.a, m->om, m->col, st.name.c_str())); break;
            case Step::UPADD: DV_TRY(op_up_dw_add(e, st.a, st.wt, st.f, st.b, st.c, st.name.c_str())); break;
            case Step::SIGMOID: DV_TRY(op_sigmoid_cols(e, maps, static_cast<long long>(N) * (H / 4) * (W / 4), 24, 2)); break;
            case Step::COPY: DV_TRY(op_copy_slice(e, st.a, st.b)); break;
            case Step::WINCONV: DV_TRY(launch_win_conv(e, st.win, st.win_flops)); break;
            case Step::DCN: DV_TRY(launch_dcn(e, st.dcn)); break;
        }
    }
    return 0;
}

// logi features of the selected cells from the resident 64-channel feature map (the last lore_detect_forward).
int lore_cell_features(Engine* e, int N, int K, int cap, const int32_t* counts, const int32_t* ax_idx, const int32_t* cr_idx,
                       float* logi_feat, int32_t* offsets_out, int32_t* overflow_host) {
    LoreNet* m = dynamic_cast<LoreNet*>(e->model.get());
    if (!m) return set_err(e, DV_ERR_STATE, "handle was not created as a lore_dla34 model");
    if (m->N != N || !m->feat.p) return set_err(e, DV_ERR_STATE, "lore_cell_features: call lore_detect_forward with the same batch first");
    if (!counts || !ax_idx || !cr_idx || !logi_feat || K <= 0 || cap <= 0) return set_err(e, DV_ERR_ARG, "lore_cell_features: bad arguments");
    if (m->K != K || m->cap != cap || m->feat_plans.empty()) DV_TRY(build_feat(e, m, K, cap));
    DV_CUDA(e, cudaMemsetAsync(m->overflow, 0, 4, e->stream));
    DV_TRY(op_cell_offsets(e, counts, N, cap, m->offsets, m->totals, m->overflow));
    if (m->resnet) DV_TRY(op_gather_pix(e, m->feat_ax, m->feat_cr, K, cap, counts, m->offsets, ax_idx, cr_idx, m->col_ax, m->col_cr));
    else DV_TRY(op_gather_patch3x3(e, m->feat, K, cap, counts, m->offsets, ax_idx, cr_idx, m->col_ax, m->col_cr));
    for (const ConvPlan& p : m->feat_plans) DV_TRY(launch_conv(e, p));
    DV_TRY(op_logi_combine(e, m->out_ax, m->out_cr, 256, cap, m->totals, logi_feat));
    if (offsets_out)
        DV_CUDA(e, cudaMemcpyAsync(offsets_out, m->offsets, (N + 1) * sizeof(int32_t), cudaMemcpyDeviceToDevice, e->stream));
    if (overflow_host) {
        DV_CUDA(e, cudaMemcpyAsync(overflow_host, m->overflow, 4, cudaMemcpyDeviceToHost, e->stream));
        DV_CUDA(e, cudaStreamSynchronize(e->stream));
    }
    return 0;
}

}  // namespace dv
