// Internal (non-ABI) declarations shared by the engine's translation units.
#pragma once
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "igemm_params.h"
#include "win_conv_params.h"
#include "mlp_params.h"
#include "dcn_params.h"

namespace dv {

// ---- error plumbing: every internal function returns 0 or a negative dv error code and leaves a
// message in Engine::err (never throws across the C ABI).
enum : int {
    DV_OK = 0,
    DV_ERR_ARG = -1,
    DV_ERR_CUDA = -2,
    DV_ERR_WEIGHTS = -3,
    DV_ERR_UNSUPPORTED = -4,
    DV_ERR_STATE = -5,
};

struct Engine;
int set_err(Engine* e, int code, const char* fmt, ...);
void set_global_err(const char* fmt, ...);

#define DV_CUDA(e, call)                                                                         \
    do {                                                                                         \
        cudaError_t _st = (call);                                                                \
        if (_st != cudaSuccess)                                                                  \
            return set_err((e), DV_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_st), \
                           __FILE__, __LINE__);                                                  \
    } while (0)
#define DV_TRY(call)              \
    do {                          \
        int _rc = (call);         \
        if (_rc != 0) return _rc; \
    } while (0)

// ---- weight blob (format written by pdf_table_b200/weights.py)
struct BlobTensor {
    void* dptr = nullptr;  // device pointer
    uint32_t dtype = 0;    // 0 f32, 1 f16, 2 i32
    uint32_t ndim = 0;
    uint32_t dims[4] = {0, 0, 0, 0};
    uint64_t nbytes = 0;
};

// NHWC fp16 activation tensor.  ld = elements between consecutive pixels (0 = dense, C); a channel slice of a
// wider (concatenation) buffer is a Tensor with p offset to its first channel and ld = the buffer's width.
struct Tensor {
    __half* p = nullptr;
    int N = 0, H = 0, W = 0, C = 0;
    int ld = 0;
    // fp32x mode: > 0 = the tensor is a split-fp16 pair, element (.., c) = hi at p[.. + c] plus lo at p[.. + c + lo]
    // (lo = C inside a [hi | lo] pixel of ld = 2C; for the padded stem image the lo copy is a second image batch)
    long long lo = 0;
    int ldc() const { return ld ? ld : C; }
    size_t elems() const { return static_cast<size_t>(N) * H * W * C; }
    Tensor slice(int coff, int c) const {
        Tensor t = *this;
        t.p = p + coff;
        t.C = c;
        t.ld = ldc();
        return t;
    }
};

struct ConvSpec {
    int KH = 1, KW = 1, stride = 1, pad = 0;
    int Cin = 0, Cout = 0;
    int BK = 64, Cin_pad = 0;  // packing of the weight matrix: [Cout][KH*KW*Cin_pad]
    const __half* w = nullptr;
    const float* bias = nullptr;  // padded to a multiple of 256 floats, or nullptr
    bool stem = false;            // 7x7 on the padded image (A_STEM): stride 2 on 4 channels, stride 1 on 8 channels
    bool flat = false;            // force A_FLAT (1x1 stride 1 / linear)
    bool split = false;           // A_FLAT split-fp16: A = [hi | lo] (2K columns), W = [W_hi | W_lo | W_hi] (3K)
};

struct EpiSpec {
    const void* res = nullptr;  // fp16, or fp32 when res_f32
    int res_mode = RES_NONE, res_ld = 0;
    int res_f32 = 0, res_mod = 0;
    int32_t* arg_out = nullptr;  // classifier arg-max epilogue (see igemm.cuh ARGMAX)
    float* max_out = nullptr;
    int act = ACT_NONE;
    int out_mode = OUT_NHWC;
    void* out = nullptr;
    int out_ld = 0, out_coff = 0, rep = 1, out_f32 = 0;
    const int* m_dyn = nullptr;  // A_FLAT: device row count (see IGemmParams::m_dyn)
    int split_off = 0;           // fp16 out: also store the fp16 residual at column + split_off
    int res_lo = 0;              // fp16 residual stored as a split pair: its lo half sits res_lo columns further
    int post_affine = 0;         // y = act(x) * post_scale + post_bias
    float post_scale = 1.f, post_bias = 0.f;
};

struct ConvPlan {
    IGemmParams prm{};
    int grid = 0;
    size_t smem = 0;
    double flops = 0;  // algorithmic 2*M*K*N (unpadded)
    double bytes = 0;  // algorithmic HBM bytes: input once + output once + weights once (+ residual)
    int res_f32 = 0;
    std::string name;
};

// One profiled launch (dv_profile_begin / dv_profile_report): CUDA events on the launching stream.
struct ProfRec {
    const char* kernel;
    std::string layer;
    double flops, bytes;
    cudaEvent_t a, b;
};

struct Model {
    virtual ~Model() {}
};

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) applies to the device that is current when it is called, so the
// "already opted in" flag of a launch site is kept per device (bit d = device d); setting the attribute twice from two
// racing threads is harmless, the flag only has to be race-free.
struct DeviceOnce {
    std::atomic<unsigned long long> done{0};
    bool need(int dev) const { return ((done.load(std::memory_order_acquire) >> (dev & 63)) & 1ull) == 0; }
    void mark(int dev) { done.fetch_or(1ull << (dev & 63), std::memory_order_release); }
};

// The dv_* entry points select the handle's device and restore the caller's current device on return.
struct DeviceGuard {
    int prev = -1;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) cudaSetDevice(dev);
        else prev = -1;
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

struct Engine {
    int device = 0;
    int num_sms = 148;
    cudaStream_t stream = nullptr;
    std::string kind;
    std::string err;
    std::map<std::string, BlobTensor> weights;
    void* weight_base = nullptr;
    std::vector<void*> owned;  // device allocations freed at destroy
    std::unique_ptr<Model> model;
    std::map<std::string, std::unique_ptr<Model>> aux;  // workspaces of the post-processing kernels, by name
    long long launches = 0;  // kernels launched through this handle (bench "gpu_launches")
    bool profiling = false;
    std::vector<ProfRec> prof;

    // Bracket every kernel launch: counts it and, when profiling, records CUDA events around it.
    void launch_begin(const char* kernel, const std::string& layer, double flops, double bytes) {
        ++launches;
        if (!profiling) return;
        ProfRec r{kernel, layer, flops, bytes, nullptr, nullptr};
        cudaEventCreate(&r.a);
        cudaEventCreate(&r.b);
        cudaEventRecord(r.a, stream);
        prof.push_back(r);
    }
    void launch_end() {
        if (profiling && !prof.empty()) cudaEventRecord(prof.back().b, stream);
    }

    int dalloc(void** p, size_t bytes, bool zero = false);
    const BlobTensor* find(const std::string& name);
};

// igemm_host.cu
int plan_conv(Engine* e, const Tensor& in, const ConvSpec& cs, const EpiSpec& es, int Ho, int Wo,
              ConvPlan* plan, const char* name);
int plan_linear(Engine* e, const __half* A, int M, int K, const ConvSpec& cs, const EpiSpec& es,
                ConvPlan* plan, const char* name, int lda = 0, int lo_off = 0);
int launch_conv(Engine* e, const ConvPlan& plan);
// conv_win_tcgen05 (win_conv.cuh): small-channel window convolution with a load/store producer
struct WinConvPlan {
    WinConvParams prm;
    int grid = 0;
    size_t smem = 0;
    double flops = 0, bytes = 0;
    std::string name;
};
int plan_win_conv(Engine* e, const Tensor& in_padded, int stride, int KR, int Ho, int Wo, const __half* w, const float* bias, int Cout,
                  int act, const Tensor& out_padded, int opad, WinConvPlan* plan, const char* name);
int launch_win_conv(Engine* e, const WinConvPlan& plan, double algorithmic_flops = 0);
bool win_patch_enabled();

// mlp_fused_tcgen05 (mlp_fused.cuh): pwconv1 -> GELU -> pwconv2 + residual (ConvNeXt) / fc1 -> GELU -> fc2 + residual (ViT)
struct MlpPlan {
    MlpParams prm;
    int C = 0, grid = 0;
    size_t smem = 0;
    double flops = 0, bytes = 0;
    std::string name;
};
bool mlp_fused_supported(int C);  // C in {96, 192, 256}; DV_MLP_FUSED=0 disables the kernel
int plan_mlp(Engine* e, const __half* h, int M, int C, const __half* w1, const float* b1, const __half* w2, const float* b2,
             float* x, MlpPlan* plan, const char* name);
int launch_mlp(Engine* e, const MlpPlan& plan);

// dcn_fused_tcgen05 (dcn_fused.cuh): deformable sampling + 3x3 GEMM + bias + ReLU, no column buffer in global memory
struct DcnPlan {
    DcnParams prm;
    int grid = 0;
    size_t smem = 0;
    double flops = 0, bytes = 0;
    std::string name;
};
bool dcn_fused_enabled();  // DV_DCN_FUSED=0 selects k_dcn_im2col + the flat GEMM
int plan_dcn(Engine* e, const Tensor& in, const float* om, const __half* w, const float* bias, int cout, int act, const Tensor& out,
             DcnPlan* plan, const char* name);
int launch_dcn(Engine* e, const DcnPlan& plan);

// ops.cu (simple HBM-bound kernels)
int op_nchw_f32_to_stem(Engine* e, const float* in, int N, int H, int W, __half* out, long long lo = 0);
int op_warp_perspective_u8(Engine* e, const uint8_t* img, int H, int W, const double* minv, const int32_t* sizes,
                           const long long* offsets, int n, int max_pixels, uint8_t* out);
int op_resize_linear_u8(Engine* e, const uint8_t* src, const long long* src_off, const int32_t* src_sizes, const int32_t* dst_widths,
                        int n, int dst_h, int dst_w_pad, uint8_t* out);
int op_crop_quads_for_rec(Engine* e, const uint8_t* pages, int H, int W, const float* quads, const int32_t* page_idx,
                          const int32_t* box_counts, int box_stride, int per_page, int n, int dst_h, int dst_w_pad, uint8_t* out,
                          int32_t* dst_widths, double* minv_ws, int32_t* sizes_ws, int width_rule = 0);
int op_warp_affine_u8(Engine* e, const uint8_t* img, int H, int W, const double* m_inv6, int w, int h, uint8_t* out);
int op_warp_affine_rects_u8(Engine* e, const uint8_t* pages, int n_pages, int H, int W, const int32_t* rects, const double* minv,
                            int n, int w, int h, uint8_t* out);
int op_pp_rec_norm(Engine* e, const uint8_t* in, const int32_t* widths, int B, int H, int W, float* out);
int op_u8_to_stem(Engine* e, const uint8_t* in, int N, int H, int W, const float* mean3, const float* std3,
                  float scale, int flip, __half* out, long long lo = 0);
int op_maxpool3x3s2(Engine* e, const Tensor& in, Tensor& out);
int op_deconv2x2_c1_sigmoid(Engine* e, const Tensor& in, const __half* w, const float* w32, float bias, float* out);
int op_nchw_f32_to_nhwc_f16(Engine* e, const float* in, int N, int C, int H, int W, __half* out);
int op_nhwc_f16_to_nchw_f32(Engine* e, const __half* in, int N, int C, int H, int W, float* out, int ld = 0, int lo = 0, int Hp = 0, int Wp = 0,
                            int py = 0, int px = 0);

// dbnet.cu
int dbnet_create(Engine* e);
int dbnet_forward(Engine* e, const float* in_nchw, const uint8_t* in_u8, const float* mean3,
                  const float* std3, float scale, int flip, int N, int H, int W, float* prob_out);
double dbnet_flops(Engine* e);
int dbnet_debug_tensor(Engine* e, const char* name, float* out_nchw, int* dims4);

// convnextvit.cu
int cnv_create(Engine* e);
int cnv_forward(Engine* e, const float* chunks, const uint8_t* crops_u8, int crop_w, int n_crops, float* logits,
                int32_t* ids, float* maxv);
int cnv_set_pass_crops(Engine* e, int crops);
int cnv_labels(Engine* e);
double cnv_flops(Engine* e);

// crnn.cu
int crnn_create(Engine* e);
int crnn_forward(Engine* e, const float* in, int N, int H, int W, float* logits, int32_t* ids, float* maxv);
int crnn_labels(Engine* e);
double crnn_flops(Engine* e);

// db_post.cu
int db_boxes(Engine* e, const float* prob, int N, int H, int W, const double* src_hw_host, float thresh, double box_thresh,
             double unclip_ratio, int max_candidates, float* boxes_out, int32_t* counts_out, int32_t* overflow_host, int variant = 0);

// lore_decode.cu
// The four small Lore head maps as strided fp32 views: element (n, c, pixel) of map m = ptr[n*img + c*chan + pixel*pix].
// Order of the stride arrays: hm (AFTER sigmoid), reg, wh, st.
struct LoreMaps {
    const float *hm = nullptr, *reg = nullptr, *wh = nullptr, *st = nullptr;
    long long img_stride[4] = {0, 0, 0, 0}, chan_stride[4] = {0, 0, 0, 0}, pix_stride[4] = {1, 1, 1, 1};
};
int lore_decode(Engine* e, const LoreMaps& maps, int N, int H, int W, const double* trans_host, int K, int MK, int wiz_rev,
                float vis_thresh, float* polygons, float* scores, int32_t* dets_feat, int32_t* ax_idx, int32_t* cr_idx,
                int32_t* counts, int32_t* rows, int32_t* overflow_host);
int centernet_decode(Engine* e, const LoreMaps& maps, int N, int H, int W, const double* trans_host, int K, int MK, float score_thr,
                     float* polygons, int32_t* counts, int32_t* overflow_host);
int lore_gather_logi(Engine* e, const float* ax, const float* cr, int N, int C, int H, int W, int K, const int32_t* counts,
                     const int32_t* ax_idx, const int32_t* cr_idx, float* logi_feat);

// lore_net.cu / lore_proc.cu
int lore_create(Engine* e);
double lore_flops(Engine* e);
int lore_debug_tensor(Engine* e, const char* name, float* out_nchw, int* dims4);
int lore_detect_forward(Engine* e, const float* in_nchw, const uint8_t* in_u8, const float* mean3, const float* std3, int flip, int N,
                        int H, int W, float* maps_out);
int lore_cell_features(Engine* e, int N, int K, int cap, const int32_t* counts, const int32_t* ax_idx, const int32_t* cr_idx,
                       float* logi_feat, int32_t* offsets_out, int32_t* overflow_host);
int lore_proc_create(Engine* e);
int lore_add_position_embeddings(Engine* e, float* feat, int cap_rows, const int32_t* dets_feat, const int32_t* counts, const int32_t* offsets,
                                 int n_img, int K);
int lore_process_forward(Engine* e, const float* feat, int cap_rows, const int32_t* rows_dev, const int32_t* offsets, int n_img,
                         float* logic_out, float* stacked_out);

// graph_net.cu (PicoDet as a graph program)
int graph_create(Engine* e);
double graph_flops(Engine* e);
int graph_num_classes(Engine* e);
int graph_debug_tensor(Engine* e, int tensor_id, float* out_nchw, int* dims4);
int picodet_forward(Engine* e, const float* in_nchw, const uint8_t* in_u8, const float* mean3, const float* std3, float scale, int flip,
                    int N, int H, int W, float* const* scores_out, float* const* dfl_out);

int rec_forward(Engine* e, const float* in_nchw, const uint8_t* in_u8, const int32_t* widths, int N, int H, int W, float* probs, int32_t* ids,
                float* maxp);
int rec_time_steps(Engine* e, int H, int W);
int ppdet_forward(Engine* e, const float* in_nchw, const uint8_t* in_u8, const float* mean3, const float* std3, float scale, int flip, int N,
                  int H, int W, float* prob_out);
int cls_forward(Engine* e, const float* in_nchw, int N, int H, int W, float* logits, float* probs);

// picodet_decode.cu
int picodet_decode(Engine* e, const float* const* scores, const float* const* dfl, int N, int C, int reg_max, const int* strides, int in_h,
                   int in_w, const float* org_hw_host, const float* scale_host, float score_thr, double iou_thr, int nms_top_k,
                   int keep_top_k, int out_cap, double* out, int32_t* counts);

// match.cu
int match_cells(Engine* e, const double* text_boxes, int n_text, const double* cell_boxes, int n_cells, int32_t* top1_out);

// ctc.cu
int ctc_collapse(Engine* e, const int32_t* ids, const float* scores, int B, int T, int blank, int32_t* out_ids,
                 int32_t* out_len, float* out_conf);
int ctc_greedy(Engine* e, const float* probs, int B, int T, int C, int blank, int32_t* out_ids,
               int32_t* out_len, float* out_conf, int32_t* raw_ids, float* raw_max);

}  // namespace dv
