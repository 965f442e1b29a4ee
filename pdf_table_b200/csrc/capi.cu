// C ABI (include/docvision.h) over the engine: handle lifetime, weight blob loading, entry points.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/docvision.h"
#include "engine.h"

namespace dv {

static thread_local std::string g_err;

void set_global_err(const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}

int set_err(Engine* e, int code, const char* fmt, ...) {
    char buf[1024];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    if (e) e->err = buf;
    g_err = buf;
    return code;
}

int Engine::dalloc(void** p, size_t bytes, bool zero) {
    if (bytes == 0) bytes = 16;
    cudaError_t st = cudaMalloc(p, bytes);
    if (st != cudaSuccess) return set_err(this, DV_ERR_CUDA, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(st));
    owned.push_back(*p);
    if (zero) {
        st = cudaMemsetAsync(*p, 0, bytes, stream);
        if (st != cudaSuccess) return set_err(this, DV_ERR_CUDA, "cudaMemset failed: %s", cudaGetErrorString(st));
    }
    return 0;
}

const BlobTensor* Engine::find(const std::string& name) {
    auto it = weights.find(name);
    return it == weights.end() ? nullptr : &it->second;
}

// ---- blob format (pdf_table_b200/weights.py): header, entry table, 256-byte aligned payload
struct BlobHeader {
    char magic[8];  // "DVWBLOB1"
    uint32_t n_tensors;
    uint32_t reserved;
    uint64_t data_offset;
    uint64_t data_bytes;
};
struct BlobEntry {
    char name[96];
    uint32_t dtype;
    uint32_t ndim;
    uint32_t dims[4];
    uint64_t offset;  // relative to data_offset
    uint64_t nbytes;
};

static int load_blob(Engine* e, const void* blob, size_t nbytes) {
    if (nbytes < sizeof(BlobHeader)) return set_err(e, DV_ERR_WEIGHTS, "weight blob too small");
    BlobHeader h;
    memcpy(&h, blob, sizeof(h));
    if (memcmp(h.magic, "DVWBLOB1", 8) != 0) return set_err(e, DV_ERR_WEIGHTS, "bad weight blob magic");
    const size_t table_end = sizeof(BlobHeader) + static_cast<size_t>(h.n_tensors) * sizeof(BlobEntry);
    if (table_end > nbytes || h.data_offset < table_end || h.data_offset + h.data_bytes > nbytes)
        return set_err(e, DV_ERR_WEIGHTS, "weight blob truncated");
    DV_CUDA(e, cudaMalloc(&e->weight_base, h.data_bytes ? h.data_bytes : 16));
    DV_CUDA(e, cudaMemcpy(e->weight_base, static_cast<const char*>(blob) + h.data_offset, h.data_bytes,
                          cudaMemcpyHostToDevice));
    const char* tab = static_cast<const char*>(blob) + sizeof(BlobHeader);
    for (uint32_t i = 0; i < h.n_tensors; ++i) {
        BlobEntry en;
        memcpy(&en, tab + static_cast<size_t>(i) * sizeof(BlobEntry), sizeof(en));
        en.name[95] = 0;
        if (en.offset + en.nbytes > h.data_bytes || (en.offset % 256) != 0)
            return set_err(e, DV_ERR_WEIGHTS, "weight blob entry '%s' out of range", en.name);
        BlobTensor t;
        t.dptr = static_cast<char*>(e->weight_base) + en.offset;
        t.dtype = en.dtype;
        t.ndim = en.ndim;
        memcpy(t.dims, en.dims, sizeof(t.dims));
        t.nbytes = en.nbytes;
        e->weights[en.name] = t;
    }
    return 0;
}

}  // namespace dv

using namespace dv;

struct dv_engine : public dv::Engine {};

extern "C" {

int dv_version(void) { return 100; }

const char* dv_last_error(dv_handle h) { return h ? h->err.c_str() : g_err.c_str(); }

int dv_create(const char* model_kind, const void* weight_blob_host, size_t nbytes, int device, dv_handle* out) {
    if (!model_kind || !out) {
        set_global_err("dv_create: null argument");
        return DV_ERR_ARG;
    }
    *out = nullptr;
    int ndev = 0;
    cudaError_t st = cudaGetDeviceCount(&ndev);
    if (st != cudaSuccess || ndev == 0) {
        set_global_err("dv_create: no CUDA device (%s) -- this engine has no CPU fallback",
                       st != cudaSuccess ? cudaGetErrorString(st) : "device count 0");
        return DV_ERR_CUDA;
    }
    if (device < 0 || device >= ndev) {
        set_global_err("dv_create: device %d out of range (%d devices)", device, ndev);
        return DV_ERR_ARG;
    }
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        set_global_err("dv_create: cannot select device %d", device);
        return DV_ERR_CUDA;
    }
    if (prop.major != 10) {
        set_global_err("dv_create: device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                       prop.minor);
        return DV_ERR_UNSUPPORTED;
    }
    dv_engine* e = new dv_engine();
    e->device = device;
    e->num_sms = prop.multiProcessorCount;
    e->kind = model_kind;
    int rc = 0;
    if (weight_blob_host && nbytes) rc = load_blob(e, weight_blob_host, nbytes);
    if (rc == 0) {
        if (e->kind == "post") {
        } else if (e->kind == "dbnet_r18") {
            rc = dbnet_create(e);
        } else if (e->kind == "crnn") {
            rc = crnn_create(e);
        } else if (e->kind == "convnext_vit") {
            rc = cnv_create(e);
        } else if (e->kind == "picodet" || e->kind == "pp_rec" || e->kind == "pplcnet_cls" || e->kind == "pp_det") {
            rc = graph_create(e);
        } else if (e->kind == "lore_dla34" || e->kind == "centernet_dla34" || e->kind == "lore_resnet18") {
            rc = lore_create(e);
        } else if (e->kind == "lore_processor") {
            rc = lore_proc_create(e);
        } else {
            rc = set_err(e, DV_ERR_UNSUPPORTED, "dv_create: unknown model kind '%s'", model_kind);
        }
    }
    if (rc != 0) {
        set_global_err("%s", e->err.c_str());
        dv_destroy(e);
        return rc;
    }
    *out = e;
    return 0;
}

int dv_destroy(dv_handle h) {
    if (!h) return 0;
    DeviceGuard dev_guard(h->device);
    cudaDeviceSynchronize();
    h->model.reset();
    h->aux.clear();
    for (void* p : h->owned) cudaFree(p);
    if (h->weight_base) cudaFree(h->weight_base);
    delete h;
    return 0;
}

int dv_set_stream(dv_handle h, void* cuda_stream) {
    if (!h) return DV_ERR_ARG;
    h->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    return 0;
}

int dv_sync(dv_handle h) {
    if (!h) return DV_ERR_ARG;
    DV_CUDA(h, cudaStreamSynchronize(h->stream));
    return 0;
}

long long dv_launch_count(dv_handle h) { return h ? h->launches : 0; }

int dv_profile_begin(dv_handle h) {
    if (!h) return DV_ERR_ARG;
    for (auto& r : h->prof) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    h->prof.clear();
    h->profiling = true;
    return 0;
}

long long dv_profile_report(dv_handle h, char* buf_host, size_t cap) {
    if (!h) return DV_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    h->profiling = false;
    DV_CUDA(h, cudaStreamSynchronize(h->stream));
    std::string js = "[";
    char line[512];
    bool first = true;
    for (auto& r : h->prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) != cudaSuccess) ms = -1.f;
        snprintf(line, sizeof(line), "%s{\"kernel\": \"%s\", \"layer\": \"%s\", \"ms\": %.6f, \"flops\": %.6e, \"bytes\": %.6e}",
                 first ? "" : ", ", r.kernel, r.layer.c_str(), ms, r.flops, r.bytes);
        js += line;
        first = false;
    }
    js += "]";
    // A buffer that is too small leaves the records in place: the caller retries with the returned size.
    if (!buf_host || js.size() + 1 > cap) return static_cast<long long>(js.size());
    memcpy(buf_host, js.data(), js.size());
    buf_host[js.size()] = 0;
    for (auto& r : h->prof) {
        cudaEventDestroy(r.a);
        cudaEventDestroy(r.b);
    }
    h->prof.clear();
    return static_cast<long long>(js.size());
}

double dv_model_flops(dv_handle h) {
    if (!h) return 0.0;
    if (h->kind == "dbnet_r18") return dbnet_flops(h);
    if (h->kind == "convnext_vit") return cnv_flops(h);
    if (h->kind == "crnn") return crnn_flops(h);
    if (h->kind == "lore_dla34" || h->kind == "centernet_dla34" || h->kind == "lore_resnet18") return lore_flops(h);
    if (h->kind == "picodet" || h->kind == "pp_rec" || h->kind == "pp_det") return graph_flops(h);
    return 0.0;
}

int dv_dbnet_forward(dv_handle h, const float* in_nchw_f32, int n, int height, int width, float* prob_out) {
    if (!h) return DV_ERR_ARG;
    if (!in_nchw_f32) return set_err(h, DV_ERR_ARG, "dv_dbnet_forward: null input");
    DeviceGuard dev_guard(h->device);
    if (h->kind == "pp_det") return ppdet_forward(h, in_nchw_f32, nullptr, nullptr, nullptr, 0.f, 0, n, height, width, prob_out);
    return dbnet_forward(h, in_nchw_f32, nullptr, nullptr, nullptr, 0.f, 0, n, height, width, prob_out);
}

int dv_dbnet_forward_u8(dv_handle h, const uint8_t* pages_hwc_u8, int n, int height, int width,
                        const float* mean3_host, const float* std3_host, float scale, int flip, float* prob_out) {
    if (!h) return DV_ERR_ARG;
    if (!pages_hwc_u8 || !mean3_host || !std3_host) return set_err(h, DV_ERR_ARG, "dv_dbnet_forward_u8: null input");
    DeviceGuard dev_guard(h->device);
    if (h->kind == "pp_det") return ppdet_forward(h, nullptr, pages_hwc_u8, mean3_host, std3_host, scale, flip, n, height, width, prob_out);
    return dbnet_forward(h, nullptr, pages_hwc_u8, mean3_host, std3_host, scale, flip, n, height, width, prob_out);
}

int dv_debug_get_tensor(dv_handle h, const char* name, float* out_nchw_f32, int* dims4_host) {
    if (!h || !name) return DV_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    if (h->kind == "dbnet_r18") return dbnet_debug_tensor(h, name, out_nchw_f32, dims4_host);
    if (h->kind == "lore_dla34" || h->kind == "centernet_dla34" || h->kind == "lore_resnet18") return lore_debug_tensor(h, name, out_nchw_f32, dims4_host);
    if ((h->kind == "picodet" || h->kind == "pp_det" || h->kind == "pp_rec") && name[0] == 't') return graph_debug_tensor(h, atoi(name + 1), out_nchw_f32, dims4_host);
    return set_err(h, DV_ERR_UNSUPPORTED, "dv_debug_get_tensor: not supported for '%s'", h->kind.c_str());
}

int dv_ctc_greedy(dv_handle h, const float* probs, int b, int t, int c, int blank, int32_t* out_ids,
                  int32_t* out_len, float* out_conf, int32_t* raw_ids, float* raw_max) {
    if (!h) return DV_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    return ctc_greedy(h, probs, b, t, c, blank, out_ids, out_len, out_conf, raw_ids, raw_max);
}

int dv_db_boxes(dv_handle h, const float* prob, int n, int height, int width, const double* src_hw_host, float thresh,
                double box_thresh, double unclip_ratio, int max_candidates, float* boxes_out, int32_t* counts_out,
                int32_t* overflow_host) {
    if (!h) return DV_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    return db_boxes(h, prob, n, height, width, src_hw_host, thresh, box_thresh, unclip_ratio, max_candidates, boxes_out,
                    counts_out, overflow_host);
}

int dv_db_boxes_dbnet(dv_handle h, const float* prob, int n, int height, int width, const double* src_hw_host, float thresh,
                      double box_thresh, double unclip_ratio, int max_candidates, float* boxes_out, int32_t* counts_out,
                      int32_t* overflow_host) {
    if (!h) return DV_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    return db_boxes(h, prob, n, height, width, src_hw_host, thresh, box_thresh, unclip_ratio, max_candidates, boxes_out,
                    counts_out, overflow_host, /*variant=*/1);
}

int dv_lore_decode(dv_handle h, const float* hm, const float* reg, const float* wh, const float* st, int layout, int n,
                   int height, int width, const double* inv_affine_host, int K, int MK, int wiz_rev, float vis_thresh,
                   float* polygons, float* scores, int32_t* dets_feat, int32_t* ax_idx, int32_t* cr_idx, int32_t* counts,
                   int32_t* rows, int32_t* overflow_host) {
    if (!h) return DV_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    LoreMaps m;
    const long long hw = static_cast<long long>(height) * width;
    if (layout == 0) {
        m.hm = hm;
        m.reg = reg;
        m.wh = wh;
        m.st = st;
        const int ch[4] = {2, 2, 8, 8};
        for (int i = 0; i < 4; ++i) {
            m.img_stride[i] = ch[i] * hw;
            m.chan_stride[i] = hw;
            m.pix_stride[i] = 1;
        }
    } else if (layout == 1) {
        if (!hm) return set_err(h, DV_ERR_ARG, "dv_lore_decode: null map");
        m.hm = hm;
        m.reg = hm + 2;
        m.wh = hm + 4;
        m.st = hm + 12;
        for (int i = 0; i < 4; ++i) {
            m.img_stride[i] = 24 * hw;
            m.chan_stride[i] = 1;
            m.pix_stride[i] = 24;
        }
    } else {
        return set_err(h, DV_ERR_ARG, "dv_lore_decode: layout must be 0 (NCHW) or 1 (NHWC x24)");
    }
    return lore_decode(h, m, n, height, width, inv_affine_host, K, MK, wiz_rev, vis_thresh, polygons, scores, dets_feat, ax_idx,
                       cr_idx, counts, rows, overflow_host);
}

static int make_lore_maps(dv_handle h, const float* hm, const float* reg, const float* wh, const float* st, int layout, int height, int width,
                          LoreMaps* m) {
    const long long hw = static_cast<long long>(height) * width;
    if (layout == 0) {
        m->hm = hm;
        m->reg = reg;
        m->wh = wh;
        m->st = st;
        const int ch[4] = {2, 2, 8, 8};
        for (int i = 0; i < 4; ++i) {
            m->img_stride[i] = ch[i] * hw;
            m->chan_stride[i] = hw;
            m->pix_stride[i] = 1;
        }
    } else if (layout == 1) {
        if (!hm) return set_err(h, DV_ERR_ARG, "null packed map");
        m->hm = hm;
        m->reg = hm + 2;
        m->wh = hm + 4;
        m->st = hm + 12;
        for (int i = 0; i < 4; ++i) {
            m->img_stride[i] = 24 * hw;
            m->chan_stride[i] = 1;
            m->pix_stride[i] = 24;
        }
    } else {
        return set_err(h, DV_ERR_ARG, "layout must be 0 (NCHW) or 1 (NHWC x24)");
    }
    return 0;
}

int dv_centernet_decode(dv_handle h, const float* hm, const float* reg, const float* c2v, const float* v2c, int layout, int n, int height,
                        int width, const double* inv_affine_host, int K, int MK, float score_threshold, float* polygons, int32_t* counts,
                        int32_t* overflow_host) {
    if (!h) return DV_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    LoreMaps m;
    DV_TRY(make_lore_maps(h, hm, reg, c2v, v2c, layout, height, width, &m));
    return centernet_decode(h, m, n, height, width, inv_affine_host, K, MK, score_threshold, polygons, counts, overflow_host);
}

int dv_centernet_forward(dv_handle h, const float* in_nchw_f32, int n, int height, int width, float* maps_out) {
    return dv_lore_detect_forward(h, in_nchw_f32, n, height, width, maps_out);
}

int dv_centernet_forward_u8(dv_handle h, const uint8_t* images_hwc_u8, int n, int height, int width, const float* mean3_host,
                            const float* std3_host, int flip, float* maps_out) {
    return dv_lore_detect_forward_u8(h, images_hwc_u8, n, height, width, mean3_host, std3_host, flip, maps_out);
}

int dv_lore_gather_logi(dv_handle h, const float* ax, const float* cr, int n, int channels, int height, int width, int K,
                        const int32_t* counts, const int32_t* ax_idx, const int32_t* cr_idx, float* logi_feat) {
    if (!h) return DV_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    return lore_gather_logi(h, ax, cr, n, channels, height, width, K, counts, ax_idx, cr_idx, logi_feat);
}

int dv_lore_detect_forward(dv_handle h, const float* in_nchw_f32, int n, int height, int width, float* maps_out) {
    if (!h) return DV_ERR_ARG;
    if (!in_nchw_f32) return set_err(h, DV_ERR_ARG, "dv_lore_detect_forward: null input");
    DeviceGuard dev_guard(h->device);
    return lore_detect_forward(h, in_nchw_f32, nullptr, nullptr, nullptr, 0, n, height, width, maps_out);
}

int dv_lore_detect_forward_u8(dv_handle h, const uint8_t* images_hwc_u8, int n, int height, int width, const float* mean3_host,
                              const float* std3_host, int flip, float* maps_out) {
    if (!h) return DV_ERR_ARG;
    if (!images_hwc_u8 || !mean3_host || !std3_host) return set_err(h, DV_ERR_ARG, "dv_lore_detect_forward_u8: null input");
    DeviceGuard dev_guard(h->device);
    return lore_detect_forward(h, nullptr, images_hwc_u8, mean3_host, std3_host, flip, n, height, width, maps_out);
}

int dv_lore_cell_features(dv_handle h, int n, int K, int max_rows, const int32_t* counts, const int32_t* ax_idx, const int32_t* cr_idx,
                          float* logi_feat, int32_t* offsets_out, int32_t* overflow_host) {
    if (!h) return DV_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    return lore_cell_features(h, n, K, max_rows, counts, ax_idx, cr_idx, logi_feat, offsets_out, overflow_host);
}

int dv_lore_add_position_embeddings(dv_handle h, float* feat, int max_rows, const int32_t* dets_feat, const int32_t* counts,
                                    const int32_t* offsets, int n_images, int K) {
    if (!h) return DV_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    return lore_add_position_embeddings(h, feat, max_rows, dets_feat, counts, offsets, n_images, K);
}

int dv_lore_process_forward(dv_handle h, const float* feat, int max_rows, const int32_t* n_rows_dev, const int32_t* offsets, int n_images,
                            float* logic_out, float* stacked_out) {
    if (!h) return DV_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    return lore_process_forward(h, feat, max_rows, n_rows_dev, offsets, n_images, logic_out, stacked_out);
}

int dv_picodet_decode(dv_handle h, const float* const* scores_host_ptrs, const float* const* dfl_host_ptrs, int n, int num_classes,
                      int reg_max, const int* strides_host, int in_height, int in_width, const float* org_hw_host,
                      const float* scale_factor_host, float score_threshold, double nms_threshold, int nms_top_k, int keep_top_k,
                      int out_cap, double* boxes_out, int32_t* counts_out) {
    if (!h) return DV_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    return picodet_decode(h, scores_host_ptrs, dfl_host_ptrs, n, num_classes, reg_max, strides_host, in_height, in_width, org_hw_host,
                          scale_factor_host, score_threshold, nms_threshold, nms_top_k, keep_top_k, out_cap, boxes_out, counts_out);
}

int dv_picodet_forward(dv_handle h, const float* in_nchw_f32, int n, int height, int width, float* const* scores_out_host_ptrs,
                       float* const* dfl_out_host_ptrs) {
    if (!h) return DV_ERR_ARG;
    if (!in_nchw_f32) return set_err(h, DV_ERR_ARG, "dv_picodet_forward: null input");
    DeviceGuard dev_guard(h->device);
    return picodet_forward(h, in_nchw_f32, nullptr, nullptr, nullptr, 1.f, 0, n, height, width, scores_out_host_ptrs, dfl_out_host_ptrs);
}

int dv_picodet_forward_u8(dv_handle h, const uint8_t* images_hwc_u8, int n, int height, int width, const float* mean3_host,
                          const float* std3_host, float scale, int flip, float* const* scores_out_host_ptrs,
                          float* const* dfl_out_host_ptrs) {
    if (!h) return DV_ERR_ARG;
    if (!images_hwc_u8 || !mean3_host || !std3_host) return set_err(h, DV_ERR_ARG, "dv_picodet_forward_u8: null input");
    DeviceGuard dev_guard(h->device);
    return picodet_forward(h, nullptr, images_hwc_u8, mean3_host, std3_host, scale, flip, n, height, width, scores_out_host_ptrs,
                           dfl_out_host_ptrs);
}

int dv_picodet_num_classes(dv_handle h) { return h ? graph_num_classes(h) : 0; }

int dv_rec_forward(dv_handle h, const float* in_nchw_f32, int n, int height, int width, float* probs_out, int32_t* ids_out, float* maxp_out) {
    if (!h) return DV_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    return rec_forward(h, in_nchw_f32, nullptr, nullptr, n, height, width, probs_out, ids_out, maxp_out);
}

int dv_rec_forward_u8(dv_handle h, const uint8_t* crops_hwc_u8, const int32_t* widths, int n, int height, int width, float* probs_out,
                      int32_t* ids_out, float* maxp_out) {
    if (!h || !crops_hwc_u8) return DV_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    return rec_forward(h, nullptr, crops_hwc_u8, widths, n, height, width, probs_out, ids_out, maxp_out);
}

int dv_cls_forward(dv_handle h, const float* in_nchw_f32, int n, int height, int width, float* logits_out, float* probs_out) {
    if (!h) return DV_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    return cls_forward(h, in_nchw_f32, n, height, width, logits_out, probs_out);
}

int dv_rec_time_steps(dv_handle h, int height, int width) { return h ? rec_time_steps(h, height, width) : 0; }
int dv_rec_num_classes(dv_handle h) { return h ? graph_num_classes(h) : 0; }

int dv_convnextvit_forward(dv_handle h, const float* chunks_nchw_f32, int n_crops, float* logits_out,
                           int32_t* ids_out, float* max_out) {
    if (!h) return DV_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    if (n_crops > 0 && !chunks_nchw_f32) return set_err(h, DV_ERR_ARG, "dv_convnextvit_forward: null input");
    return cnv_forward(h, chunks_nchw_f32, nullptr, 0, n_crops, logits_out, ids_out, max_out);
}

int dv_convnextvit_forward_u8(dv_handle h, const uint8_t* crops_hwc_u8, int n_crops, int crop_w, float* logits_out,
                              int32_t* ids_out, float* max_out) {
    if (!h) return DV_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    if (n_crops > 0 && !crops_hwc_u8) return set_err(h, DV_ERR_ARG, "dv_convnextvit_forward_u8: null input");
    return cnv_forward(h, nullptr, crops_hwc_u8, crop_w, n_crops, logits_out, ids_out, max_out);
}

int dv_convnextvit_labels(dv_handle h) { return h ? cnv_labels(h) : 0; }

int dv_crnn_forward(dv_handle h, const float* in_nchw_f32, int n, int height, int width, float* logits_out, int32_t* ids_out,
                    float* max_out) {
    if (!h) return DV_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    if (n > 0 && (!in_nchw_f32 || !ids_out)) return set_err(h, DV_ERR_ARG, "dv_crnn_forward: null input / output");
    return crnn_forward(h, in_nchw_f32, n, height, width, logits_out, ids_out, max_out);
}

int dv_crnn_labels(dv_handle h) { return h ? crnn_labels(h) : 0; }

int dv_match_cells(dv_handle h, const double* text_boxes, int n_text, const double* cell_boxes, int n_cells, int32_t* top1_out) {
    if (!h) return DV_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    return match_cells(h, text_boxes, n_text, cell_boxes, n_cells, top1_out);
}

int dv_convnextvit_set_pass_crops(dv_handle h, int crops) {
    if (!h) return DV_ERR_ARG;
    return cnv_set_pass_crops(h, crops);
}

int dv_warp_perspective_u8(dv_handle h, const uint8_t* page_hwc_u8, int height, int width, const double* minv, const int32_t* sizes,
                           const int64_t* offsets, int n, int max_pixels, uint8_t* out) {
    if (!h) return DV_ERR_ARG;
    if (n == 0) return 0;
    if (!page_hwc_u8 || !minv || !sizes || !offsets || !out || n < 0 || height <= 0 || width <= 0 || max_pixels <= 0)
        return set_err(h, DV_ERR_ARG, "dv_warp_perspective_u8: null pointer / bad size");
    DeviceGuard dev_guard(h->device);
    return op_warp_perspective_u8(h, page_hwc_u8, height, width, minv, sizes, reinterpret_cast<const long long*>(offsets), n, max_pixels, out);
}

int dv_resize_linear_u8(dv_handle h, const uint8_t* src_packed, const int64_t* src_offsets, const int32_t* src_sizes,
                        const int32_t* dst_widths, int n, int dst_h, int dst_w_pad, uint8_t* out) {
    if (!h) return DV_ERR_ARG;
    if (n == 0) return 0;
    if (!src_packed || !src_offsets || !src_sizes || !dst_widths || !out || n < 0 || dst_h <= 0 || dst_w_pad <= 0)
        return set_err(h, DV_ERR_ARG, "dv_resize_linear_u8: null pointer / bad size");
    DeviceGuard dev_guard(h->device);
    return op_resize_linear_u8(h, src_packed, reinterpret_cast<const long long*>(src_offsets), src_sizes, dst_widths, n, dst_h, dst_w_pad, out);
}

int dv_crop_quads_for_rec(dv_handle h, const uint8_t* pages_hwc_u8, int n_pages, int height, int width, const float* quads,
                          const int32_t* page_idx, int n, int dst_h, int dst_w_pad, uint8_t* out, int32_t* dst_widths, double* minv_ws,
                          int32_t* sizes_ws, int width_rule) {
    if (!h) return DV_ERR_ARG;
    if (n == 0) return 0;
    if (!pages_hwc_u8 || !quads || !out || !dst_widths || !minv_ws || !sizes_ws || n < 0 || n_pages <= 0 || height <= 0 || width <= 0 ||
        dst_h <= 0 || dst_w_pad <= 0)
        return set_err(h, DV_ERR_ARG, "dv_crop_quads_for_rec: null pointer / bad size");
    DeviceGuard dev_guard(h->device);
    return op_crop_quads_for_rec(h, pages_hwc_u8, height, width, quads, page_idx, nullptr, 0, 0, n, dst_h, dst_w_pad, out, dst_widths, minv_ws,
                                 sizes_ws, width_rule);
}

int dv_crop_boxes_for_rec(dv_handle h, const uint8_t* pages_hwc_u8, int n_pages, int height, int width, const float* boxes,
                          const int32_t* box_counts, int box_stride, int per_page, int dst_h, int dst_w_pad, uint8_t* out,
                          int32_t* dst_widths, double* minv_ws, int32_t* sizes_ws, int width_rule) {
    if (!h) return DV_ERR_ARG;
    if (!pages_hwc_u8 || !boxes || !box_counts || !out || !dst_widths || !minv_ws || !sizes_ws || n_pages <= 0 || height <= 0 || width <= 0 ||
        box_stride <= 0 || per_page <= 0 || per_page > box_stride || dst_h <= 0 || dst_w_pad <= 0)
        return set_err(h, DV_ERR_ARG, "dv_crop_boxes_for_rec: null pointer / bad size");
    DeviceGuard dev_guard(h->device);
    return op_crop_quads_for_rec(h, pages_hwc_u8, height, width, boxes, nullptr, box_counts, box_stride, per_page, n_pages * per_page, dst_h,
                                 dst_w_pad, out, dst_widths, minv_ws, sizes_ws, width_rule);
}

int dv_warp_affine_u8(dv_handle h, const uint8_t* img_hwc_u8, int height, int width, const double* m_inv6_host, int out_w, int out_h,
                      uint8_t* out) {
    if (!h) return DV_ERR_ARG;
    if (!img_hwc_u8 || !m_inv6_host || !out || height <= 0 || width <= 0 || out_w <= 0 || out_h <= 0)
        return set_err(h, DV_ERR_ARG, "dv_warp_affine_u8: null pointer / bad size");
    if (static_cast<long long>(out_w) * out_h > 0x7fffffffLL / 4) return set_err(h, DV_ERR_ARG, "dv_warp_affine_u8: output too large");
    DeviceGuard dev_guard(h->device);
    return op_warp_affine_u8(h, img_hwc_u8, height, width, m_inv6_host, out_w, out_h, out);
}

int dv_crop_tables_for_tsr(dv_handle h, const uint8_t* pages_hwc_u8, int n_pages, int height, int width, const int32_t* rects,
                           const double* m_inv, int n, int out_w, int out_h, uint8_t* out) {
    if (!h) return DV_ERR_ARG;
    if (n < 0 || n > 65535 || n_pages <= 0 || height <= 0 || width <= 0 || out_w <= 0 || out_h <= 0)
        return set_err(h, DV_ERR_ARG, "dv_crop_tables_for_tsr: bad size (0 <= n <= 65535)");
    if (n > 0 && (!pages_hwc_u8 || !rects || !m_inv || !out)) return set_err(h, DV_ERR_ARG, "dv_crop_tables_for_tsr: null pointer");
    if (static_cast<long long>(out_w) * out_h > 0x7fffffffLL / 4) return set_err(h, DV_ERR_ARG, "dv_crop_tables_for_tsr: output too large");
    DeviceGuard dev_guard(h->device);
    return op_warp_affine_rects_u8(h, pages_hwc_u8, n_pages, height, width, rects, m_inv, n, out_w, out_h, out);
}

int dv_pp_rec_normalise(dv_handle h, const uint8_t* crops_hwc_u8, const int32_t* widths, int b, int height, int width,
                        float* out_nchw_f32) {
    if (!h) return DV_ERR_ARG;
    if (!crops_hwc_u8 || !widths || !out_nchw_f32 || b <= 0 || height <= 0 || width <= 0)
        return set_err(h, DV_ERR_ARG, "dv_pp_rec_normalise: null pointer / empty batch");
    DeviceGuard dev_guard(h->device);
    return op_pp_rec_norm(h, crops_hwc_u8, widths, b, height, width, out_nchw_f32);
}

int dv_ctc_collapse(dv_handle h, const int32_t* ids, const float* scores, int b, int t, int blank,
                    int32_t* out_ids, int32_t* out_len, float* out_conf) {
    if (!h) return DV_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    return ctc_collapse(h, ids, scores, b, t, blank, out_ids, out_len, out_conf);
}

int dv_conv2d_nhwc_f16(dv_handle h, const void* in_nhwc_f16, int n, int height, int width, int cin,
                       const void* weight_packed_f16, int cin_pad, const float* bias, int cout, int ksize,
                       int stride, int pad, const void* residual_nhwc_f16, int act, void* out_nhwc_f16) {
    if (!h) return DV_ERR_ARG;
    if (!in_nhwc_f16 || !weight_packed_f16 || !out_nhwc_f16) return set_err(h, DV_ERR_ARG, "dv_conv2d: null pointer");
    DeviceGuard dev_guard(h->device);
    Tensor in;
    in.p = const_cast<__half*>(reinterpret_cast<const __half*>(in_nhwc_f16));
    in.N = n;
    in.H = height;
    in.W = width;
    in.C = cin;
    const int Ho = (height + 2 * pad - ksize) / stride + 1;
    const int Wo = (width + 2 * pad - ksize) / stride + 1;
    ConvSpec cs;
    cs.KH = cs.KW = ksize;
    cs.stride = stride;
    cs.pad = pad;
    cs.Cin = cin;
    cs.Cout = cout;
    cs.Cin_pad = cin_pad;
    cs.BK = (cin_pad % 64 == 0) ? 64 : (cin_pad % 32 == 0) ? 32 : 16;
    cs.w = reinterpret_cast<const __half*>(weight_packed_f16);
    cs.bias = bias;
    EpiSpec es;
    es.out = out_nhwc_f16;
    es.out_ld = cout;
    es.act = act;
    if (residual_nhwc_f16) {
        es.res = reinterpret_cast<const __half*>(residual_nhwc_f16);
        es.res_mode = RES_SAME;
        es.res_ld = cout;
    }
    ConvPlan plan;
    const size_t owned_before = h->owned.size();
    int rc = plan_conv(h, in, cs, es, Ho, Wo, &plan, "dv_conv2d");
    if (rc == 0) rc = launch_conv(h, plan);
    // one-shot plan: release its delta table once the kernel has run
    cudaError_t sync_st = cudaStreamSynchronize(h->stream);
    if (rc == 0 && sync_st != cudaSuccess) rc = set_err(h, DV_ERR_CUDA, "dv_conv2d: %s", cudaGetErrorString(sync_st));
    while (h->owned.size() > owned_before) {
        cudaFree(h->owned.back());
        h->owned.pop_back();
    }
    if (rc == 0) {
        cudaError_t st = cudaGetLastError();
        if (st != cudaSuccess) rc = set_err(h, DV_ERR_CUDA, "dv_conv2d: %s", cudaGetErrorString(st));
    }
    return rc;
}

int dv_nchw_f32_to_nhwc_f16(dv_handle h, const float* in, int n, int c, int height, int width, void* out) {
    if (!h) return DV_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    return op_nchw_f32_to_nhwc_f16(h, in, n, c, height, width, reinterpret_cast<__half*>(out));
}
int dv_nhwc_f16_to_nchw_f32(dv_handle h, const void* in, int n, int c, int height, int width, float* out) {
    if (!h) return DV_ERR_ARG;
    DeviceGuard dev_guard(h->device);
    return op_nhwc_f16_to_nchw_f32(h, reinterpret_cast<const __half*>(in), n, c, height, width, out);
}

}  // extern "C"
