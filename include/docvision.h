/*
 * docvision.h -- C ABI of libdocvision.so, the B200 (sm_100a) engine that replaces the per-page
 * vision hot path of CycloneBoy/pdf_table behind the reference's predictor API.
 *
 * Reference interface this ABI stands in for (paths relative to the reference repo,
 * src/pdftable/model/ocr_pdf/):
 *   - BaseInferTask.infer / infer_pytorch / build_*_infer_batch   base_infer_task.py:317-381
 *     (the "trt" predictor slot, _prepare_trt_mode :143-144, is the slot a maintainer fills, see
 *     INTEGRATION.md)
 *   - OcrDetectionTask._run_model / _postprocess                  ocr_detection_task.py:89-141
 *   - OcrRecognitionTask._run_model / _postprocess                ocr_recognition_task.py:81-136
 *
 * Conventions
 *   - Every function returns 0 on success or a negative dv_status; nothing throws across the ABI.
 *     dv_last_error(h) returns a thread-unsafe, handle-owned message (h == NULL: creation errors).
 *   - All data pointers are DEVICE pointers owned by the caller unless the name ends in `_host`.
 *   - Work is enqueued asynchronously on the handle's stream (dv_set_stream); call dv_sync or
 *     synchronise the stream yourself before reading results.
 *   - A handle is bound to one device, is not thread-safe, and distinct handles are independent.
 *   - No allocation happens on the hot path once a (batch, H, W) shape has been seen: activation
 *     workspaces and TMA descriptors are planned on first use of a shape and reused.
 */
#ifndef DOCVISION_H_
#define DOCVISION_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct dv_engine* dv_handle;

enum dv_status {
    DV_STATUS_OK = 0,
    DV_STATUS_BAD_ARGUMENT = -1,
    DV_STATUS_CUDA_ERROR = -2,
    DV_STATUS_BAD_WEIGHTS = -3,
    DV_STATUS_UNSUPPORTED = -4,
    DV_STATUS_BAD_STATE = -5
};

/* activation / residual / store modes of dv_conv2d_nhwc_f16 (mirror csrc/igemm.cuh) */
enum dv_act { DV_ACT_NONE = 0, DV_ACT_RELU = 1, DV_ACT_GELU = 2, DV_ACT_SIGMOID = 3, DV_ACT_HSWISH = 4 };

int dv_version(void);
const char* dv_last_error(dv_handle h);

/*
 * Create an engine handle.
 *   model_kind : "post"          post-processing kernels only (no weights)
 *                "dbnet_r18"     DBNet ResNet-18 text detector   (reference model/db_net/dbnet.py:715-728)
 *                "convnext_vit"  ConvNextViT text-line recogniser (reference model/convnext_vit/
 *                                modeling_convnext_vit.py:20-45)
 *                "lore_dla34"    Lore table-structure detector, DLA-34 + DCNv2 (reference model/lore/lore_dla_34.py:193)
 *                "lore_resnet18" Lore `wireless` key-point detector, ResNet-18 + transposed-conv up path
 *                                (reference model/lore/lore_detector.py:148-389)
 *                "centernet_dla34" CenterNet table-structure detector (reference model/center_net/modeling_centernet.py:601)
 *                "picodet"       PicoDet layout detector as a graph program (reference model/picodet/{lcnet,csp_pan,pico_head}.py)
 *                "lore_processor" Lore logical-location transformers (reference model/lore/lore_processor.py:399)
 *   weight_blob: HOST pointer to a blob written by pdf_table_b200.weights.pack_* (may be NULL for "post")
 * Replaces: BaseInferTask._get_inference_model / DeployUtils.model_eval (base_infer_task.py:146-169,
 * utils/deploy_utils.py:226-240).
 */
int dv_create(const char* model_kind, const void* weight_blob_host, size_t nbytes, int device, dv_handle* out);
int dv_destroy(dv_handle h);
int dv_set_stream(dv_handle h, void* cuda_stream);
int dv_sync(dv_handle h);
/* number of kernels launched through this handle so far (bench.py "gpu_launches") */
long long dv_launch_count(dv_handle h);
/* algorithmic FLOPs of one forward at the last planned shape (0 if none) */
double dv_model_flops(dv_handle h);
/*
 * Per-launch device timing (bench.py "roofline"): between dv_profile_begin and dv_profile_report every
 * kernel launched through the handle is bracketed by CUDA events on the handle's stream.
 * dv_profile_report synchronises, stops profiling and writes a JSON array
 *   [{"kernel": "...", "layer": "...", "ms": t, "flops": F, "bytes": B}, ...]   (one object per launch)
 * into buf_host (capacity cap bytes, NUL-terminated).  Returns the number of bytes required (excluding
 * the NUL; call again with a larger buffer if >= cap) or a negative dv_status.  No reference counterpart
 * (the reference only wall-clocks whole stages, ocr_system_task.py:646-660).
 */
int dv_profile_begin(dv_handle h);
long long dv_profile_report(dv_handle h, char* buf_host, size_t cap);

/*
 * DB text-detector forward: fp32 NCHW pages -> fp32 probability map [N,1,H,W].  The handle's model kind selects the
 * network: "dbnet_r18" (the in-tree DBModel, db_net/dbnet.py:715) or "pp_det" (the PP-OCRv4 mobile detector, PPLCNetV3-0.75 +
 * RSE-FPN + DBHead, the graph the reference downloads as ONNX: ocr_table_model_config.py:134-147).
 * Replaces OcrDetectionTask._run_model (ocr_detection_task.py:89-124): predictor(image) -> pred[0].
 * H and W must be multiples of 32 (DetResizeForTest guarantees it, db_pp/image_operators.py:304-305).
 */
int dv_dbnet_forward(dv_handle h, const float* in_nchw_f32, int n, int height, int width, float* prob_out);
/*
 * Same network fed with raw uint8 HWC pages; fuses PPOcrDetectionPreprocessor's channel flip +
 * NormalizeImage + ToCHWImage (db_pp/processor_ocr_db_pp.py:124, image_operators.py:93-118):
 *   x[c] = (u8[flip ? 2-c : c] * scale - mean[c]) / std[c]     (fp32, then fp16 NHWC)
 * mean3/std3 are HOST pointers.
 */
int dv_dbnet_forward_u8(dv_handle h, const uint8_t* pages_hwc_u8, int n, int height, int width,
                        const float* mean3_host, const float* std3_host, float scale, int flip,
                        float* prob_out);

/*
 * CTC greedy decode of a [B,T,C] fp32 probability tensor.
 * Replaces CTCLabelDecode.__call__ + BaseRecLabelDecode.decode(is_remove_duplicate=True)
 * (ocr_rec_pp/rec_postprocess.py:175-191, 126-161).
 *   out_ids  [B,T] int32 : kept class ids, left-packed, padded with -1
 *   out_len  [B]   int32 : number of kept ids
 *   out_conf [B]   fp32  : np.mean of the kept per-step maxima (0 when nothing is kept)
 *   raw_ids / raw_max [B,T] : optional (may be NULL) per-step argmax / max before collapsing
 * The id -> character lookup (self.character[text_id]) stays on the host.
 */
int dv_ctc_greedy(dv_handle h, const float* probs, int b, int t, int c, int blank, int32_t* out_ids,
                  int32_t* out_len, float* out_conf, int32_t* raw_ids, float* raw_max);

/*
 * DB "threshold + box seed": probability maps -> text boxes, entirely on the device.
 * Replaces OcrDetectionTask._postprocess (ocr_detection_task.py:126-141) = PPOcrDetectionPostProcessor.__call__
 * (db_pp/processor_ocr_db_pp.py:330-342): DBPostProcess.__call__ / boxes_from_bitmap / get_mini_boxes /
 * box_score_fast / unclip (:174-311) followed by filter_tag_det_res (:374-386).
 *   prob          : [n,1,height,width] fp32 (device), the output of dv_dbnet_forward
 *   src_hw_host   : HOST [n][2] doubles (src_h, src_w) = shape_list[:2] = org_shape[:2] of each page
 *   thresh        : binarisation threshold (compared in fp32, as numpy does); CLI default 0.2
 *   box_thresh    : minimum mean probability inside the box (0.6); unclip_ratio (1.5); max_candidates (<= 1000)
 *   boxes_out     : [n][max_candidates][8] fp32 (device) -- x0,y0,..,x3,y3 clockwise from top-left, source-image pixels,
 *                   in the reference's contour order, left-packed
 *   counts_out    : [n] int32 (device) boxes per page
 *   overflow_host : HOST int32 or NULL; receives the number of contours skipped because they exceed the kernel's
 *                   vertex capacity (2048 vertices after CHAIN_APPROX_SIMPLE); non-NULL makes the call synchronous
 */
int dv_db_boxes(dv_handle h, const float* prob, int n, int height, int width, const double* src_hw_host, float thresh,
                double box_thresh, double unclip_ratio, int max_candidates, float* boxes_out, int32_t* counts_out,
                int32_t* overflow_host);
/*
 * The same for the reference's in-tree DBNet back-end (model="db"): replaces OCRDetectionPostProcessor.__call__
 * (db_net/processor_ocr_dbnet.py:113-127) = boxes_from_bitmap (db_net/ocr_detection_utils.py:168-205).  Differences from
 * dv_db_boxes: the reference's constants there are box_thresh 0.3 / unclip 1.5 / 1000 contours (pass them); the mini box is
 * truncated to int32 BEFORE it is scaled (np.round(box / width * dest_width), float64) and clipped to [0, dest]; corners keep
 * get_mini_boxes' order (top-left, top-right, bottom-right, bottom-left of the x-sorted pairs) and no
 * filter_tag_det_res step follows.  src_hw_host = org_shape (height, width) of each page.
 */
int dv_db_boxes_dbnet(dv_handle h, const float* prob, int n, int height, int width, const double* src_hw_host, float thresh,
                      double box_thresh, double unclip_ratio, int max_candidates, float* boxes_out, int32_t* counts_out,
                      int32_t* overflow_host);

/*
 * Lore / CenterNet "heat-map 3x3 max-pool NMS + top-K gather": head maps -> sorted table cells, on the device.
 * Replaces process_detect_output (lore/lineless_table_process.py:592-655) called from LoreModel.forward
 * (lore/modeling_lore.py:146-152): corner_decode :97-124, ctdet_4ps_decode incl. the wiz_rev corner snapping
 * :127-267, ctdet_4ps_post_process :489-507, merge_outputs / filter / normalized_ps :551-589.
 *   hm, reg, wh, st : fp32 head maps (device).  layout 0: four NCHW tensors [n,2,h,w], [n,2,h,w], [n,8,h,w], [n,8,h,w];
 *                     layout 1: one NHWC tensor [n,h,w,24] passed as `hm` (channels hm0,hm1,reg0,reg1,wh0-7,st0-7,pad4;
 *                     reg/wh/st ignored).  `hm` is the map AFTER the sigmoid (the reference's first step, :599).
 *   inv_affine_host : HOST [n][6] doubles, the row-major 2x3 matrix get_affine_transform(c, s, 0, (out_w, out_h), inv=1)
 *                     (:403-438) built from the reference's int64 meta (lore/processer_lore.py:112-130)
 *   K, MK           : top-K cells / corners (reference 3000 / 5000); wiz_rev, vis_thresh: LoreConfig (wtw: 1, 0.2)
 *   polygons  [n][K][8] fp32 : cell corners in source pixels, rows in the reference's (re-sorted) order = results[1][:, :8]
 *   scores    [n][K]    fp32 : penalised scores = results[1][:, 8]
 *   dets_feat [n][K][8] int32: slct_dets_feat (feature-map corner coordinates truncated and clamped to 0..255)
 *   ax_idx [n][K], cr_idx [n][K][4] int32 : flat h*w indices the logical feature of row j is gathered from
 *                     (`ax` at the cell centre, `cr` at the four cc_match corners -- dv_lore_gather_logi)
 *   counts [n] int32 : rows with score >= vis_thresh (= num_valid); rows [n] int32 or NULL: rows written per image.
 *                     Rows are written for every peak above min(0.2, vis_thresh); the reference's remaining top-K
 *                     padding (NMS-suppressed zeros at arbitrary positions) is never selected and is not produced.
 *   overflow_host   : HOST int32 or NULL; non-zero if an image had more than 16384 gated peaks of one class
 */
int dv_lore_decode(dv_handle h, const float* hm, const float* reg, const float* wh, const float* st, int layout, int n,
                   int height, int width, const double* inv_affine_host, int K, int MK, int wiz_rev, float vis_thresh,
                   float* polygons, float* scores, int32_t* dets_feat, int32_t* ax_idx, int32_t* cr_idx, int32_t* counts,
                   int32_t* rows, int32_t* overflow_host);

/*
 * Logical-location features of the selected cells from DENSE `ax` / `cr` maps ([n,channels,h,w] fp32, NCHW):
 * logi_feat[i][j] = ax[:, ax_idx[j]] + sum_k cr[:, cr_idx[j][k]] for j < counts[i]   ([n][K][channels] fp32).
 * Replaces _tranpose_and_gather_feat(ax) + _get_4ps_feat(cc_match, cr).sum(3) + `logi + cr`
 * (lore/lineless_table_process.py:31-63, 148, 253-254, 644).  The network path computes the same rows without ever
 * materialising the dense 512-channel maps (dv_lore_forward).
 */
int dv_lore_gather_logi(dv_handle h, const float* ax, const float* cr, int n, int channels, int height, int width, int K,
                        const int32_t* counts, const int32_t* ax_idx, const int32_t* cr_idx, float* logi_feat);

/*
 * Lore table-structure detector (model kind "lore_dla34"): DLA-34 + DCNv2 neck + the four small heads.
 * Replaces LoreModel.forward's `self.detect_infer_model(pixel_values)` (lore/modeling_lore.py:143) =
 * DLASeg.forward (lore/lore_dla_34.py:176-190) for get_dla_dcn(34, {'hm':2,'st':8,'wh':8,'ax':256,'cr':256,'reg':2}).
 *   in_nchw_f32 : [n,3,height,width] fp32, the output of TableLorePreProcessor (lore/processer_lore.py:66-109);
 *                 height, width multiples of 32 (1024 x 1024 for the wtw configuration)
 *   _u8 variant : images_hwc_u8 [n,height,width,3] = the warpAffine output; ((x / 255. - mean) / std) is evaluated in
 *                 float64 as numpy does (:92) and fused into the stem's input kernel; flip swaps channels 0 and 2
 *   maps_out    : [n,height/4,width/4,24] fp32 NHWC or NULL (kept inside the handle): hm0,hm1 (AFTER sigmoid), reg0,reg1,
 *                 wh0-7, st0-7, 4 pad -- the `layout 1` input of dv_lore_decode.
 * The 256-channel `ax` / `cr` heads are NOT evaluated densely (the reference only gathers them at the selected
 * cells): the 64-channel feature map stays resident in the handle for dv_lore_cell_features.
 * Model kind "lore_resnet18" (the `wireless` configuration, LoreDetectModel.forward lore/lore_detector.py:353-389) takes the
 * same calls and returns the same packed map; height, width multiples of 64 (768 x 768 for the configuration); its `ax` /
 * `cr` heads are evaluated densely up to their last 64-channel hidden maps, which stay resident for dv_lore_cell_features
 * (the final 1x1 convs run only at the selected cells).
 */
int dv_lore_detect_forward(dv_handle h, const float* in_nchw_f32, int n, int height, int width, float* maps_out);
int dv_lore_detect_forward_u8(dv_handle h, const uint8_t* images_hwc_u8, int n, int height, int width, const float* mean3_host,
                              const float* std3_host, int flip, float* maps_out);

/*
 * Logical-location features of the selected cells, computed from the feature map of the last dv_lore_detect_forward
 * on this handle: the `ax` head (3x3 conv + ReLU + 1x1) at each cell centre plus the `cr` head at its four cc_match
 * corners.  Same values as gathering the reference's dense maps (dv_lore_gather_logi), at 5 x cells rows instead of
 * height/4 x width/4 pixels.  Replaces the `ax` / `cr` heads of DLASeg.forward together with
 * _tranpose_and_gather_feat / _get_4ps_feat (lore/lineless_table_process.py:31-63).  On a "lore_resnet18" handle the two heads'
 * four 3x3 convs were evaluated densely by dv_lore_detect_forward (their 64-channel hidden maps are resident) and this call runs
 * the last 1x1 conv (64 -> 256) on the gathered pixels only.
 *   counts [n], ax_idx [n][K], cr_idx [n][K][4] : outputs of dv_lore_decode (device)
 *   max_rows     : capacity of the packed row list (sum of counts over the batch must fit)
 *   logi_feat    : [max_rows][256] fp32 (device); image i's cells occupy rows offsets[i] .. offsets[i+1]
 *   offsets_out  : [n+1] int32 (device) or NULL; offsets_out[n] = total rows (the n_rows_dev of dv_lore_process_forward)
 *   overflow_host: HOST int32 or NULL; receives the total when it exceeds max_rows (non-NULL makes the call synchronous)
 */
int dv_lore_cell_features(dv_handle h, int n, int K, int max_rows, const int32_t* counts, const int32_t* ax_idx, const int32_t* cr_idx,
                          float* logi_feat, int32_t* offsets_out, int32_t* overflow_host);

/*
 * Lore logical-location processor (model kind "lore_processor").
 * Replaces LoreProcessModel.forward, evaluation branch without 2-D position embeddings (wtw)
 * (lore/lore_processor.py:465-514) called from LoreModel.forward (lore/modeling_lore.py:158-166).
 *   feat       : [max_rows][256] fp32 (device) cell features; n_rows_dev: device int32, number of valid rows
 *   offsets    : [n_images+1] int32 (device): attention runs inside each image's row segment
 *   logic_out  : [max_rows][4] fp32 or NULL (base regressor); stacked_out: [max_rows][4] fp32 (stacking regressor =
 *                the `logits` the task rounds with process_logic_output)
 */
int dv_lore_process_forward(dv_handle h, const float* feat, int max_rows, const int32_t* n_rows_dev, const int32_t* offsets, int n_images,
                            float* logic_out, float* stacked_out);
/*
 * The 2-D position embeddings of the wiz_2dpe configurations ("ptn", "wireless": lore/configuration_lore.py:72-116), added to the
 * cell features before dv_lore_process_forward: feat += x_pe[d0] + y_pe[d1] + x_pe[d2] + y_pe[d5] with d = the integer position
 * features of the decode (LoreProcessModel.forward, lore/lore_processor.py:486-490; called with dets=slct_dets_feat from
 * LoreModel.forward, lore/modeling_lore.py:155-159).  In place on a "lore_processor" handle (its blob holds the two tables).
 *   feat [max_rows][256] fp32 (device); dets_feat [n_images][K][8] int32, counts [n_images], offsets [n_images+1] (device):
 *   outputs of dv_lore_decode / dv_lore_cell_features.
 */
int dv_lore_add_position_embeddings(dv_handle h, float* feat, int max_rows, const int32_t* dets_feat, const int32_t* counts,
                                    const int32_t* offsets, int n_images, int K);

/*
 * CenterNet table-structure detector (model kind "centernet_dla34"): DLA-34 + plain IDA-up + heads hm / v2c / c2v / reg.
 * Replaces OcrTableStructureTask._run_model for model="CenterNet" (ocr_table_structure_task.py:176-204) =
 * DLASeg.forward (center_net/modeling_centernet.py:657-662).  Same calling convention as dv_lore_detect_forward; maps_out is
 * the packed [n,height/4,width/4,24] fp32 map with channels hm0,hm1 (after sigmoid), reg0,reg1, c2v0-7, v2c0-7, 4 pad.
 */
int dv_centernet_forward(dv_handle h, const float* in_nchw_f32, int n, int height, int width, float* maps_out);
int dv_centernet_forward_u8(dv_handle h, const uint8_t* images_hwc_u8, int n, int height, int width, const float* mean3_host,
                            const float* std3_host, int flip, float* maps_out);

/*
 * CenterNet decode: head maps -> table-cell polygons, on the device.
 * Replaces OCRTableCenterNetPostProcessor.__call__ (center_net/processer_centernet.py:170-204): bbox_decode / gbox_decode
 * (center_net/table_process.py:151-216: 3x3 max-pool NMS + top-K + gathers), bbox_post_process / gbox_post_process (:219-236),
 * group_bbox_by_gbox (:278-333: cell corners snap to detected vertices), the score > threshold filter and the sort by
 * 0.01 * mean_x + mean_y.  (table_process.nms :239-275 is a no-op in the reference: it receives the [1,K,10] batch array.)
 *   hm, reg, c2v, v2c : layout 0 = four NCHW fp32 tensors [n,2|2|8|8,h,w]; layout 1 = the packed NHWC x24 map as `hm`
 *   inv_affine_host   : HOST [n][6] doubles, get_affine_transform(c, s, 0, (out_w, out_h), inv=1) with the float meta
 *   K, MK             : top-K cells / vertices (reference 1000 / 4000); score_threshold 0.3
 *   polygons [n][K][8] fp32 (device), rows in the reference's final order; counts [n] int32
 */
int dv_centernet_decode(dv_handle h, const float* hm, const float* reg, const float* c2v, const float* v2c, int layout, int n, int height,
                        int width, const double* inv_affine_host, int K, int MK, float score_threshold, float* polygons, int32_t* counts,
                        int32_t* overflow_host);

/*
 * PicoDet layout detector (model kind "picodet"): LCNet-x1.0 backbone + CSP-PAN neck + PicoHead, executed from the graph
 * program that pdf_table_b200.picodet_graph lowers the reference modules to (picodet/lcnet.py:159-263,
 * picodet/csp_pan.py:233-360, picodet/pico_head.py:37-167, 1108-1138).
 * Replaces OcrLayoutTask._run_model (ocr_layout_task.py:84-123: the ONNX session of picodet_lcnet_x1_0_fgd_layout*) with
 * the output contract of PicoHead.forward_eval(export_post_process=False) (pico_head.py:1130-1138).
 *   in_nchw_f32 : [n,3,height,width] fp32, the output of OCRPicodetPreProcessor (picodet/processor_picodet.py:72-113); 800 x 608
 *   _u8 variant : images_hwc_u8 [n,height,width,3] = the cv2.resize output; the channel flip (:94) and
 *                 (x * scale - mean) / std in fp32 (:66-70) are fused into the stem kernel
 *   scores_out_host_ptrs[4] : HOST array of 4 DEVICE pointers, level l: fp32 [n, HW_l, num_classes] sigmoid class scores
 *   dfl_out_host_ptrs[4]    : HOST array of 4 DEVICE pointers, level l: fp32 [n, HW_l, 32] raw DFL logits
 *                             (HW_l = ceil(height/s_l) * ceil(width/s_l), s = 8, 16, 32, 64) -- the inputs of dv_picodet_decode
 */
int dv_picodet_forward(dv_handle h, const float* in_nchw_f32, int n, int height, int width, float* const* scores_out_host_ptrs,
                       float* const* dfl_out_host_ptrs);
int dv_picodet_forward_u8(dv_handle h, const uint8_t* images_hwc_u8, int n, int height, int width, const float* mean3_host,
                          const float* std3_host, float scale, int flip, float* const* scores_out_host_ptrs,
                          float* const* dfl_out_host_ptrs);
int dv_picodet_num_classes(dv_handle h);

/*
 * PP-OCR text-line recogniser (SURVEY.md a5, 8(b) `dv_rec_forward`): the PP-OCRv4 "SVTR-LCNet" network -- PPLCNetV3-0.95
 * backbone, SVTR neck (2 global-mixer blocks, 8 heads), CTC head -- for model kind "pp_rec" (weights packed by
 * pdf_table_b200/pp_rec_graph.py).  Replaces the ONNX session run of OcrRecognitionTask._run_model for
 * model="PP-OCRv4" (ocr_pdf/ocr_recognition_task.py:84-116, the hub model of ocr_pdf/ocr_table_model_config.py:166-204)
 * together with the arg-max / max half of CTCLabelDecode.__call__ (ocr_rec_pp/rec_postprocess.py:175-183).
 *   in_nchw_f32 : device fp32 [n,3,height,width] = PPOcrRecPreProcessor's batch (height 48; a4 / dv_pp_rec_normalise)
 *   probs_out   : device fp32 [n,T,C] softmax probabilities, or NULL (the [n,T,C] tensor is then never written)
 *   ids_out     : device int32 [n,T] per-step arg-max, or NULL;   maxp_out : device fp32 [n,T] per-step max probability, or NULL
 *                 (feed both to dv_ctc_collapse for the decoded ids and the mean confidence)
 *   T = dv_rec_time_steps(h, height, width) (= width / 8 for widths that are multiples of 8), C = dv_rec_num_classes(h).
 * dv_rec_forward_u8: uint8 HWC crops [n,height,width,3] already resized to `height`, left aligned, valid up to widths[n]
 * (device int32 [n], NULL = full width); (x / 255 - 0.5) / 0.5 and the zero padding beyond each width are fused into the stem.
 */
int dv_rec_forward(dv_handle h, const float* in_nchw_f32, int n, int height, int width, float* probs_out, int32_t* ids_out, float* maxp_out);
int dv_rec_forward_u8(dv_handle h, const uint8_t* crops_hwc_u8, const int32_t* widths, int n, int height, int width, float* probs_out,
                      int32_t* ids_out, float* maxp_out);
int dv_rec_time_steps(dv_handle h, int height, int width);
/*
 * PULC image classifiers: PP-LCNet x1.0 (cls/cls_pp_lcnet.py:164-293) for model kind "pplcnet_cls" (weights
 * packed by pdf_table_b200/pplcnet_graph.py; the stride list of the task -- (2,1) strides for textline_orientation /
 * language_classification, cls/configuration_cls_pulc.py:20-42 -- is part of the packed program).  Replaces the torch forward of
 * ClsImagePulcTask._run_model (ocr_pdf/cls_image_pulc_task.py:61-83).
 *   in_nchw_f32 : device fp32 [n,3,height,width] = PPLCNetImageProcessor's pixel_values
 *   logits_out  : device fp32 [n,C] = PPLCNet.forward's return value (what TableAttribute thresholds), or NULL
 *   probs_out   : device fp32 [n,C] = softmax(logits) (what Topk sorts), or NULL;   C = dv_rec_num_classes(h)
 */
int dv_cls_forward(dv_handle h, const float* in_nchw_f32, int n, int height, int width, float* logits_out, float* probs_out);
int dv_rec_num_classes(dv_handle h);

/*
 * PicoDet "anchor decode": head outputs of the four FPN levels -> layout boxes, on the device.
 * Replaces OcrLayoutTask._postprocess (ocr_layout_task.py:125-157) = OCRPicodetPostProcessor.__call__
 * (picodet/processor_picodet.py:184-298) with hard_nms / iou_of / area_of (:301-360) and warp_boxes (:136-158).
 *   scores_host_ptrs[4] : HOST array of 4 DEVICE pointers, level l: fp32 [n, HW_l, num_classes] sigmoid class scores
 *   dfl_host_ptrs[4]    : HOST array of 4 DEVICE pointers, level l: fp32 [n, HW_l, 4*(reg_max+1)] raw DFL logits
 *                         (the `export_post_process=False` contract of PicoHead.forward_eval, picodet/pico_head.py:1130-1138);
 *                         HW_l = ceil(in_height/stride_l) * ceil(in_width/stride_l)
 *   org_hw_host [n][2]  : original (height, width) of each page; scale_factor_host [n][2] = (ratio_h, ratio_w) of the resize
 *   score_threshold 0.5, nms_threshold 0.5, nms_top_k 1000 (<= 1000), keep_top_k 100 (<= 100): PicodetConfig defaults
 *   boxes_out [n][out_cap][6] float64 (device): rows (class id, score, x1, y1, x2, y2) in the reference's order (classes
 *                         ascending, NMS pick order inside a class), original-image pixels; counts_out [n] int32.
 * float32 / float64 exactly where the reference has them; softmax uses expf (numpy's SIMD exp may differ in the last ulp).
 */
int dv_picodet_decode(dv_handle h, const float* const* scores_host_ptrs, const float* const* dfl_host_ptrs, int n, int num_classes,
                      int reg_max, const int* strides_host, int in_height, int in_width, const float* org_hw_host,
                      const float* scale_factor_host, float score_threshold, double nms_threshold, int nms_top_k, int keep_top_k,
                      int out_cap, double* boxes_out, int32_t* counts_out);

/*
 * ConvNextViT text-line recogniser forward.
 * Replaces OcrRecognitionTask._run_model for model="ConvNextViT" (ocr_recognition_task.py:81-116) =
 * OCRRecognition.forward (ocr_recognition/modeling_ocr_recognition.py:137-149) -> ConvNextViT.forward
 * (convnext_vit/modeling_convnext_vit.py:37-45), fused with the arg-max of OCRRecognitionPostProcessor
 * (ocr_recognition/processor_ocr_recognition.py:147-151; softmax is monotone, so arg-max of the logits).
 *   chunks_nchw_f32 : [3*n_crops, 3, 32, 300] fp32 in [0,1], the output of OCRRecognitionPreprocessor (:73-115)
 *   logits_out      : [n_crops, 201, L] fp32 or NULL (parity dump; L = dv_convnextvit_labels)
 *   ids_out         : [n_crops, 201] int32 per-token arg-max (torch.argmax: first maximum)
 *   max_out         : [n_crops, 201] fp32 maximum logit or NULL
 */
int dv_convnextvit_forward(dv_handle h, const float* chunks_nchw_f32, int n_crops, float* logits_out,
                           int32_t* ids_out, float* max_out);
/*
 * Same network fed with uint8 HWC crops [n_crops, 32, crop_w, 3] (height already 32, zero padded on the right to
 * a common crop_w <= 804): fuses the pad-to-804 / 3-chunk split (x0 = 0, 252, 504; width 300) / `/255` / NHWC->NCHW of
 * OCRRecognitionPreprocessor.__call__ (ocr_recognition/processor_ocr_recognition.py:57-61, 104-112) and the
 * RGB->gray of ConvNextViT.forward into the patchify kernel.  The keep-ratio cv2.resize (:44-56) stays with the caller.
 */
int dv_convnextvit_forward_u8(dv_handle h, const uint8_t* crops_hwc_u8, int n_crops, int crop_w, float* logits_out,
                              int32_t* ids_out, float* max_out);
int dv_convnextvit_labels(dv_handle h);

/*
 * CRNN text-line recogniser forward (model kind "crnn").
 * Replaces OcrRecognitionTask._run_model for model="CRNN" (ocr_recognition_task.py:81-116) = OCRRecognition.forward
 * (ocr_recognition/modeling_ocr_recognition.py:137-149) -> CRNN.forward (crnn/modeling_crnn.py:90-113: RGB -> gray, seven
 * convs + BatchNorm + ReLU with max-pools, the (2,1) row-folding conv, two bidirectional LSTMs + Linear, 512 -> L classifier),
 * fused with the arg-max of OCRRecognitionPostProcessor (ocr_recognition/processor_ocr_recognition.py:147-151).
 *   in_nchw_f32 : [n, 3, 32, width] fp32 in [0,1] (the output of OCRRecognitionPreprocessor, :73-115; width % 4 == 0)
 *   logits_out  : [n, width / 4, L] fp32 or NULL (parity dump; L = dv_crnn_labels)
 *   ids_out     : [n, width / 4] int32 per-step arg-max (torch.argmax: first maximum); collapse with dv_ctc_collapse(blank 0)
 *   max_out     : [n, width / 4] fp32 maximum logit or NULL
 */
int dv_crnn_forward(dv_handle h, const float* in_nchw_f32, int n, int height, int width, float* logits_out, int32_t* ids_out,
                    float* max_out);
int dv_crnn_labels(dv_handle h);
/* crops per internal pass (default 96): sizes the activation workspace so the widest tensor stays near L2 */
int dv_convnextvit_set_pass_crops(dv_handle h, int crops);
/*
 * det -> rec glue: the perspective crops of OcrCommonUtils.crop_image
 * (utils/ocr/ocr_common_utils.py:214-262), i.e. cv2.warpPerspective(page, T, (w, h)) with its defaults (INTER_LINEAR,
 * BORDER_CONSTANT 0), n crops of one page in one launch, bit-exact against cv2 (OpenCV 4.13 fixed-point remap).
 * page_hwc_u8: device uint8 [height, width, 3].  minv: device double [n, 9] = cv2.invert(T) per crop (the host keeps
 * getPerspectiveTransform / invert, as the reference).  sizes: device int32 [n, 2] = (w, h) of each crop.  offsets:
 * device int64 [n] byte offset of crop i in `out` (h*w*3 bytes each, HWC).  max_pixels = max w*h.
 */
int dv_warp_perspective_u8(dv_handle h, const uint8_t* page_hwc_u8, int height, int width, const double* minv, const int32_t* sizes,
                           const int64_t* offsets, int n, int max_pixels, uint8_t* out);
/*
 * cv2.resize(crop, (dst_widths[i], dst_h)) with the default INTER_LINEAR on uint8 HWC crops, bit-exact against cv2 (OpenCV 4.13:
 * 11-bit coefficients, the 2x2 box average for an exact 2x reduction).  Replaces the resize of
 * OCRRecognitionPreprocessor.keepratio_resize (ocr_recognition/processor_ocr_recognition.py:44-62) for crops that are
 * already on the device.  src_packed / src_offsets (int64 [n], bytes) / src_sizes (int32 [n,2] = (w, h)): the packed crops as
 * dv_warp_perspective_u8 writes them.  out: device uint8 [n, dst_h, dst_w_pad, 3], columns >= dst_widths[i] zero -- the layout
 * dv_convnextvit_forward_u8 reads.
 */
int dv_resize_linear_u8(dv_handle h, const uint8_t* src_packed, const int64_t* src_offsets, const int32_t* src_sizes,
                        const int32_t* dst_widths, int n, int dst_h, int dst_w_pad, uint8_t* out);
/*
 * The whole det -> rec glue on the device, no host round trip: for each detected quad
 * OcrCommonUtils.crop_image (utils/ocr/ocr_common_utils.py:214-262: corner ordering, crop size, cv2.getPerspectiveTransform,
 * cv2.warpPerspective) followed by OCRRecognitionPreprocessor.keepratio_resize (processor_ocr_recognition.py:44-62), bit-exact
 * against the cv2 calls of the reference (OpenCV 4.13 arithmetic restated: float products + LU with partial pivoting for the
 * homography, closed-form 3x3 inverse, fixed-point remap and resize).  pages: device uint8 [n_pages, height, width, 3];
 * quads: device float32 [n, 4, 2] in any corner order (dv_db_boxes output); page_idx: device int32 [n] or NULL (all page 0).
 * out: device uint8 [n, dst_h, dst_w_pad, 3] = the input of dv_convnextvit_forward_u8 (dst_h 32, dst_w_pad 804);
 * dst_widths: device int32 [n], 0 where the reference's cv2 call would raise (empty crop / singular quad; the row block is
 * zero).  minv_ws (double [n, 9]) and sizes_ws (int32 [n, 2] = crop (w, h)) are caller-owned workspaces, readable afterwards.
 * width_rule 0 = that keep-ratio rule (ConvNextViT: dst_h 32, dst_w_pad 804); width_rule 1 = the PP-OCR recogniser's
 * resize_norm_img for a crop that is its own batch, as the reference's orchestrator calls it (ocr_rec_pp/processor_ocr_rec_pp.py:
 * 43-59; dst_h 48, dst_w_pad 1280): dst_widths[i] = min(imgW, max(ceil(48 w / h), 16)), imgW = clamp(int(48 max(w / h, 320 / 48)),
 * 16, 1280) being the padded width the host recomputes from sizes_ws to group the crops for dv_rec_forward_u8.
 */
int dv_crop_quads_for_rec(dv_handle h, const uint8_t* pages_hwc_u8, int n_pages, int height, int width, const float* quads,
                          const int32_t* page_idx, int n, int dst_h, int dst_w_pad, uint8_t* out, int32_t* dst_widths, double* minv_ws,
                          int32_t* sizes_ws, int width_rule);
/*
 * The same, reading dv_db_boxes' outputs directly: boxes device float32 [n_pages, box_stride, 8], box_counts device int32
 * [n_pages]; every page gets per_page crop slots (slot k = box k of that page, skipped with width 0 when k >= its count), so
 * the detector's result never visits the host between detection and recognition.  out: [n_pages * per_page, dst_h, dst_w_pad, 3].
 */
int dv_crop_boxes_for_rec(dv_handle h, const uint8_t* pages_hwc_u8, int n_pages, int height, int width, const float* boxes,
                          const int32_t* box_counts, int box_stride, int per_page, int dst_h, int dst_w_pad, uint8_t* out,
                          int32_t* dst_widths, double* minv_ws, int32_t* sizes_ws, int width_rule);
/*
 * cv2.warpAffine(img, M, (out_w, out_h), flags=INTER_LINEAR) with a zero border on a uint8 HWC image, bit-exact against cv2
 * (OpenCV 4.13 fixed point): the warp of TableLorePreProcessor.process (lore/processer_lore.py:80-91) for an
 * image that is already on the device.  m_inv6_host: HOST pointer to the six doubles of the INVERTED matrix (cv2 inverts M
 * first; predictors.invert_affine is that formula).  out: device uint8 [out_h, out_w, 3].
 */
int dv_warp_affine_u8(dv_handle h, const uint8_t* img_hwc_u8, int height, int width, const double* m_inv6_host, int out_w, int out_h,
                      uint8_t* out);
/*
 * Layout -> table-structure glue: the table loop of the orchestrator (ocr_pdf/ocr_system_task.py:184-198),
 * i.e. OcrCommonUtils.crop_image_by_box (utils/ocr/ocr_common_utils.py:269-284) followed by the cv2.warpAffine of
 * TableLorePreProcessor.process (lore/processer_lore.py:80-91), for all tables of a batch of resident pages in one launch and
 * without the JPEG write / re-read of the reference (the device path sees the page's exact pixels).
 *   pages_hwc_u8 : device uint8 [n_pages][height][width][3]
 *   rects        : device int32 [n][5] = page, x0, y0, crop_w, crop_h -- the slice img[y0:y0+crop_h, x0:x0+crop_w]
 *   m_inv        : device double [n][6], the INVERTED 2x3 matrix of each crop (as for dv_warp_affine_u8)
 *   out          : device uint8 [n][out_h][out_w][3] = cv2.warpAffine(crop, M, (out_w, out_h), INTER_LINEAR), bit-exact; a rect
 *                  that is not inside its page gives a zero image.  This is the batch dv_lore_detect_forward_u8 reads.
 */
int dv_crop_tables_for_tsr(dv_handle h, const uint8_t* pages_hwc_u8, int n_pages, int height, int width, const int32_t* rects,
                           const double* m_inv, int n, int out_w, int out_h, uint8_t* out);
/*
 * PP-OCR recogniser pre-process after the host cv2.resize: replaces the numpy tail of
 * PPOcrRecPreProcessor.resize_norm_img (ocr_rec_pp/processor_ocr_rec_pp.py:56-63): astype(float32), HWC -> CHW, / 255,
 * -= 0.5, /= 0.5 and the zero padding to the batch width.  crops_hwc_u8: device uint8 [b, height, width, 3], crop i
 * left-aligned with widths[i] valid columns (device int32 [b]); out: device fp32 [b, 3, height, width].  Bit-exact.
 */
int dv_pp_rec_normalise(dv_handle h, const uint8_t* crops_hwc_u8, const int32_t* widths, int b, int height, int width,
                        float* out_nchw_f32);
/*
 * Cell / text matching of the table export: for every recognised text box the table cell it belongs to.
 * Replaces the loop of OcrTableToHtmlTask.match_table_cell_and_text_cell over find_top1_mach_box
 * (ocr_pdf/ocr_table_to_html_task.py:48-77, 196-207): the first cell (list order) that contains the text box
 * (box_in_other_box with diff 2, pdf_table/table_common.py:138-160), else the cell with the smallest
 * (1 - compute_iou_v2, distance) key, first occurrence on ties (table_common.py:435-441, 473-516).
 *   text_boxes : device float64 [n_text, 4] = x1, y1, x2, y2 (OcrCell.to_bbox);  cell_boxes : device float64 [n_cells, 4] (Cell.to_bbox)
 *   top1_out   : device int32 [n_text] cell index per text box.  float64 arithmetic in the reference's operation order: identical indices.
 */
int dv_match_cells(dv_handle h, const double* text_boxes, int n_text, const double* cell_boxes, int n_cells, int32_t* top1_out);
/*
 * Collapse per-step arg-max ids: keep[t] = ids[t] != blank && (t == 0 || ids[t] != ids[t-1]).
 * Replaces the loop of OCRRecognitionPostProcessor.__call__ (processor_ocr_recognition.py:152-162) and, given
 * scores, the confidence of BaseRecLabelDecode.decode (ocr_rec_pp/rec_postprocess.py:126-161).
 * Outputs as dv_ctc_greedy; scores may be NULL (out_conf = 0).
 */
int dv_ctc_collapse(dv_handle h, const int32_t* ids, const float* scores, int b, int t, int blank,
                    int32_t* out_ids, int32_t* out_len, float* out_conf);

/*
 * Operator-level entry used by the parity tests of the tensor-core convolution kernel:
 * NHWC fp16 convolution (stride 1 or 2, square kernel, zero padding) with fused bias, optional residual
 * add and activation.  weight_packed: fp16 [cout][kh*kw*cin_pad] (see weights.pack_conv), bias: fp32
 * padded to a multiple of 256 (or NULL), residual: NHWC fp16 [n,ho,wo,cout] or NULL.
 * Replaces torch.nn.functional.conv2d as used by every nn.Conv2d on the path (e.g. dbnet.py:23-31).
 */
int dv_conv2d_nhwc_f16(dv_handle h, const void* in_nhwc_f16, int n, int height, int width, int cin,
                       const void* weight_packed_f16, int cin_pad, const float* bias, int cout, int ksize,
                       int stride, int pad, const void* residual_nhwc_f16, int act, void* out_nhwc_f16);
/*
 * Debug / parity aid: copy a named intermediate activation of the last forward (e.g. "c1", "layer2.0",
 * "fuse", "b2" for dbnet_r18) as fp32 NCHW.  dims4_host (HOST, may be NULL) receives {N,C,H,W};
 * out_nchw_f32 (DEVICE) may be NULL to query the shape only.  No reference counterpart.
 */
int dv_debug_get_tensor(dv_handle h, const char* name, float* out_nchw_f32, int* dims4_host);
/* layout helpers for the tests */
int dv_nchw_f32_to_nhwc_f16(dv_handle h, const float* in, int n, int c, int height, int width, void* out);
int dv_nhwc_f16_to_nchw_f32(dv_handle h, const void* in, int n, int c, int height, int width, float* out);

#ifdef __cplusplus
}
#endif
#endif /* DOCVISION_H_ */
