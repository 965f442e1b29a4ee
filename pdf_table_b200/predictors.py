"""Host-side mirror of the reference's predictor API for the hot path.

Same class names, constructor arguments, call shape (construct -> _preprocess -> _run_model -> _postprocess,
``__call__ = post(run(pre(x)))``), return conventions and error behaviour as
``pdftable.model.ocr_pdf.{base_infer_task, ocr_detection_task, ocr_recognition_task}`` -- so a maintainer can
register these behind ``predictor_type="b200"`` and the orchestrator keeps working.  Everything
numeric runs in libdocvision.so on the GPU; there is no CPU / PyTorch fallback: constructing a task without the
library or without a B200 raises.

What stays on the host, exactly as in the reference: decoding the input (path / PIL / ndarray), ``cv2.resize`` to
the network size (DetResizeForTest, keepratio_resize) and the id -> character lookup.
"""
from __future__ import annotations

import math
import os
from typing import Any, Dict, List, Mapping, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import weights
from ._lib import DocVisionError
from .engine import Engine

__all__ = ["BaseInferTask", "OcrDetectionTask", "OcrRecognitionTask", "OcrTableStructureTask", "OcrLayoutTask", "ClsImagePulcTask", "det_resize_for_test",
           "keepratio_resize", "lore_affine", "lore_affine_upper_left", "lore_preprocess", "pp_rec_batch_plan", "PPOcrRecPreProcessor", "crop_geometry", "crop_images", "crops_for_recognition", "invert_affine", "lore_preprocess_device", "pp_rec_padded_width", "pp_rec_launch_groups", "table_crop_rect", "sort_det_boxes", "order_point", "order_points_batch", "det_resize_shape", "det_resize_for_test_device", "dbnet_resize_shape"]


def _read_image(inputs) -> np.ndarray:
    """Input decoding of the reference processors (db_pp/processor_ocr_db_pp.py:113-122): path, PIL or ndarray."""
    if isinstance(inputs, str):
        from PIL import Image

        return np.array(Image.open(inputs).convert("RGB"))
    try:
        import PIL.Image

        if isinstance(inputs, PIL.Image.Image):
            return np.array(inputs)
    except ImportError:  # pragma: no cover
        pass
    if isinstance(inputs, np.ndarray):
        return inputs
    raise TypeError(f"inputs should be either str, PIL.Image, np.array, but got {type(inputs)}")


def det_resize_for_test(img: np.ndarray, limit_side_len: int = 960, limit_type: str = "max"):
    """DetResizeForTest.resize_image_type0 (db_pp/image_operators.py:269-316): scale so the max (min) side meets the
    limit, round each side to a multiple of 32 (at least 32), cv2.resize (bilinear).  Returns (img, [ratio_h, ratio_w])."""
    import cv2

    h, w = img.shape[:2]
    if h + w < 64:  # DetResizeForTest.__call__ / image_padding (image_operators.py:236-239, 254-258)
        pad = np.zeros((max(32, h), max(32, w), img.shape[2]), np.uint8)
        pad[:h, :w, :] = img
        img = pad
        h, w = img.shape[:2]
    resize_h, resize_w = det_resize_shape(h, w, limit_side_len, limit_type)
    if (resize_h, resize_w) != (h, w):
        img = cv2.resize(img, (int(resize_w), int(resize_h)))
    return img, [resize_h / float(h), resize_w / float(w)]


def det_resize_shape(h: int, w: int, limit_side_len: int = 960, limit_type: str = "max"):
    """The size rule of DetResizeForTest.resize_image_type0 (db_pp/image_operators.py:283-305) -> (resize_h, resize_w)."""
    if limit_type == "max":
        ratio = float(limit_side_len) / max(h, w) if max(h, w) > limit_side_len else 1.0
    elif limit_type == "min":
        ratio = float(limit_side_len) / min(h, w) if min(h, w) < limit_side_len else 1.0
    elif limit_type == "resize_long":
        ratio = float(limit_side_len) / max(h, w)
    else:
        raise Exception("not support limit type, image ")
    resize_h = max(int(round(int(h * ratio) / 32) * 32), 32)
    resize_w = max(int(round(int(w * ratio) / 32) * 32), 32)
    return resize_h, resize_w


def det_resize_for_test_device(engine: Engine, page: torch.Tensor, limit_side_len: int = 960, limit_type: str = "max"):
    """det_resize_for_test for a page that is already on the device (uint8 HWC cuda tensor): the same size rule, the
    cv2.resize by ``dv_resize_linear_u8`` (bit-exact against cv2).  Returns (uint8 cuda tensor, [ratio_h, ratio_w])."""
    h, w = int(page.shape[0]), int(page.shape[1])
    if h + w < 64:
        pad = torch.zeros((max(32, h), max(32, w), page.shape[2]), dtype=torch.uint8, device=page.device)
        pad[:h, :w, :] = page
        page = pad
        h, w = int(page.shape[0]), int(page.shape[1])
    resize_h, resize_w = det_resize_shape(h, w, limit_side_len, limit_type)
    if (resize_h, resize_w) != (h, w):
        page = engine.resize_pages_u8(page.unsqueeze(0), resize_w, resize_h)[0]
    return page, [resize_h / float(h), resize_w / float(w)]


def keepratio_resize(img: np.ndarray, target_height: int = 32, target_width: int = 804) -> np.ndarray:
    """OCRRecognitionPreprocessor.keepratio_resize (ocr_recognition/processor_ocr_recognition.py:44-62) WITHOUT the
    zero padding to target_width (the engine pads): returns uint8 [32, cur_w <= 804, 3]."""
    import cv2

    cur_ratio = img.shape[1] / float(img.shape[0])
    if cur_ratio > float(target_width) / target_height:
        cur_w = target_width
    else:
        cur_w = int(target_height * cur_ratio)
    return cv2.resize(img, (cur_w, target_height))


def sort_det_boxes(det_result: np.ndarray) -> np.ndarray:
    """The reading-order sort the orchestrator applies to the detector's boxes before recognition
    (ocr_pdf/ocr_system_task.py:158-162): key = 0.01 * mean x + mean y on the python floats of ``tolist()``, stable."""
    rows = np.asarray(det_result).tolist()
    rows = sorted(rows, key=lambda x: 0.01 * sum(x[::2]) / 4 + sum(x[1::2]) / 4)
    return np.array(rows)


def order_point(coor) -> np.ndarray:
    """OcrCommonUtils.order_point (utils/ocr/ocr_common_utils.py:287-305): the four corners sorted by their angle around the
    centroid (numpy's default argsort), rotated so that the first one lies left of the centroid; float32 [4,2]."""
    arr = np.array(coor).reshape([4, 2])
    centroid = np.sum(arr, 0) / arr.shape[0]
    theta = np.arctan2(arr[:, 1] - centroid[1], arr[:, 0] - centroid[0])
    pts = arr[np.argsort(theta)].reshape([4, -1])
    if pts[0][0] > centroid[0]:
        pts = np.concatenate([pts[3:], pts[:3]])
    return pts.reshape([4, 2]).astype("float32")


def crop_geometry(position):
    """The host part of OcrCommonUtils.crop_image (utils/ocr/ocr_common_utils.py:227-257): order the four corners (by x,
    then the left pair and the right pair by y), take the crop size from the distances between opposite edge mid-points
    (python float maths), and build the float32 source / destination corner arrays of the homography.
    Returns (corners [4,2] float32, corners_trans [4,2] float32, (w, h))."""
    import math

    position = np.asarray(position).tolist()
    for i in range(4):
        for j in range(i + 1, 4):
            if position[i][0] > position[j][0]:
                position[i], position[j] = position[j], position[i]
    if position[0][1] > position[1][1]:
        position[0], position[1] = position[1], position[0]
    if position[2][1] > position[3][1]:
        position[2], position[3] = position[3], position[2]
    (x1, y1), (x4, y4), (x2, y2), (x3, y3) = position[0], position[1], position[2], position[3]
    corners = np.zeros((4, 2), np.float32)
    corners[0], corners[1], corners[2], corners[3] = [x1, y1], [x2, y2], [x4, y4], [x3, y3]
    img_width = math.sqrt(pow((x1 + x4) / 2 - (x2 + x3) / 2, 2) + pow((y1 + y4) / 2 - (y2 + y3) / 2, 2))
    img_height = math.sqrt(pow((x1 + x2) / 2 - (x4 + x3) / 2, 2) + pow((y1 + y2) / 2 - (y4 + y3) / 2, 2))
    trans = np.zeros((4, 2), np.float32)
    trans[1], trans[2], trans[3] = [img_width - 1, 0], [0, img_height - 1], [img_width - 1, img_height - 1]
    return corners, trans, (int(img_width), int(img_height))


def crop_images(engine: Engine, page: torch.Tensor, positions) -> List[Optional[torch.Tensor]]:
    """OcrCommonUtils.crop_image for every detected quad of one page, on the GPU: the homography is
    solved on the host with cv2.getPerspectiveTransform exactly as the reference does (and inverted with cv2.invert, which
    is what cv2.warpPerspective does first); the warp itself -- all crops of the page -- is one launch of
    ``dv_warp_perspective_u8``, bit-exact against cv2.  page: uint8 [H,W,3] cuda.  Returns one uint8 [h,w,3] cuda tensor per
    quad, or None where the reference's cv2 call would raise (a crop of zero width or height)."""
    import cv2

    minv, sizes, keep = [], [], []
    for k, pos in enumerate(positions):
        corners, trans, (w, h) = crop_geometry(pos)
        if w <= 0 or h <= 0:
            continue
        minv.append(cv2.invert(cv2.getPerspectiveTransform(corners, trans))[1])
        sizes.append((w, h))
        keep.append(k)
    out: List[Optional[torch.Tensor]] = [None] * len(positions)
    if keep:
        for k, crop in zip(keep, engine.warp_perspective_u8(page, np.stack(minv), np.array(sizes, np.int32))):
            out[k] = crop
    return out


def crops_for_recognition(engine: Engine, page: torch.Tensor, positions, target_height: int = 32, target_width: int = 804):
    """det -> rec on the device: OcrCommonUtils.crop_image for every quad (``crop_images``) followed by
    OCRRecognitionPreprocessor.keepratio_resize (ocr_recognition/processor_ocr_recognition.py:44-62) of every crop, both
    bit-exact against cv2, without the crops ever visiting the host.  Returns (crops uint8 [n, 32, Wc, 3] cuda, zero padded
    to the widest crop -- the input of ``Engine.convnextvit_forward_u8`` --, widths [n], kept quad indices); quads whose crop
    or resized width is empty (the reference's cv2 calls raise there and the orchestrator skips the crop) are dropped."""
    import cv2

    minv, sizes, widths, keep = [], [], [], []
    for k, pos in enumerate(positions):
        corners, trans, (w, h) = crop_geometry(pos)
        if w <= 0 or h <= 0:
            continue
        cur_ratio = w / float(h)
        cur_w = target_width if cur_ratio > float(target_width) / target_height else int(target_height * cur_ratio)
        if cur_w <= 0:
            continue
        minv.append(cv2.invert(cv2.getPerspectiveTransform(corners, trans))[1])
        sizes.append((w, h))
        widths.append(cur_w)
        keep.append(k)
    if not keep:
        return torch.empty((0, target_height, 0, 3), dtype=torch.uint8, device=page.device), [], []
    _, packed = engine.warp_perspective_u8(page, np.stack(minv), np.array(sizes, np.int32), return_packed=True)
    out = engine.resize_linear_u8(packed, np.array(widths, np.int32), target_height, max(widths))
    return out, widths, keep


def pp_rec_batch_plan(shapes, rec_image_shape=(3, 48, 320), rec_batch_num: int = 6, limited_min_width: int = 16,
                      limited_max_width: int = 1280):
    """The host rule of PPOcrRecPreProcessor (ocr_rec_pp/processor_ocr_rec_pp.py:44-59, 100-121): crops sorted by aspect
    ratio (``np.argsort``, as the reference), batches of ``rec_batch_num``; per batch W = int(48 * max(max ratio, 320/48))
    clamped to [min, max]; per crop resized_w = min(W, max(ceil(48 * ratio), min)).  shapes: [(h, w)].
    Returns (indices, [(batch_beg_img_no, W, [resized_w ...])])."""
    import math

    _, img_h, img_w0 = rec_image_shape
    indices = np.argsort(np.array([w / float(h) for h, w in shapes]))
    plan = []
    for beg in range(0, len(shapes), rec_batch_num):
        ids = indices[beg:beg + rec_batch_num]
        max_wh_ratio = 0
        for i in ids:
            h, w = shapes[i]
            max_wh_ratio = max(max_wh_ratio, w * 1.0 / h)
        max_wh_ratio = max(max_wh_ratio, img_w0 / img_h)
        img_w = max(min(int(img_h * max_wh_ratio), limited_max_width), limited_min_width)
        widths = []
        for i in ids:
            h, w = shapes[i]
            ratio_w = max(math.ceil(img_h * (w / float(h))), limited_min_width)
            widths.append(img_w if ratio_w > img_w else int(ratio_w))
        plan.append((beg, img_w, widths))
    return indices, plan


def pp_rec_padded_width(sizes, img_h: int = 48, img_w0: int = 320, limited_min_width: int = 16, limited_max_width: int = 1280) -> np.ndarray:
    """The padded width PPOcrRecPreProcessor.resize_norm_img gives a crop that is its own batch (processor_ocr_rec_pp.py:43-49),
    vectorised over crop sizes [n,2] = (w, h): imgW = clamp(int(48 * max(w / h, 320 / 48)), 16, 1280); 0 for an empty crop."""
    sizes = np.asarray(sizes).reshape(-1, 2)
    w, h = sizes[:, 0].astype(np.float64), sizes[:, 1].astype(np.float64)
    ok = (w > 0) & (h > 0)
    ratio = np.where(ok, w * 1.0 / np.where(ok, h, 1.0), 0.0)
    img_w = (img_h * np.maximum(ratio, img_w0 / img_h)).astype(np.int64)
    return np.where(ok, np.clip(img_w, limited_min_width, limited_max_width), 0).astype(np.int64)


def pp_rec_launch_groups(rec: Engine, post: Engine, crops: torch.Tensor, widths: torch.Tensor, sizes: torch.Tensor, t_max: int = 160):
    """PP-OCR recogniser over the output of dv_crop_quads_for_rec / dv_crop_boxes_for_rec with width_rule 1: crops uint8
    [n,48,1280,3], widths int32 [n] (resized width, 0 = skipped quad), sizes int32 [n,2] (all cuda).  The reference recognises one
    crop per call, i.e. every crop padded to ITS OWN width imgW (``pp_rec_padded_width``); the network is per-crop, so crops of
    equal imgW share a launch.  The crop sizes come back to the host once (8 bytes per crop) to form the groups.  Returns
    (ids int32 [n,t_max] left-packed / -1 padded, lens int32 [n], conf fp32 [n]) on the device."""
    n = int(crops.shape[0])
    dev = crops.device
    TRANSFER["d2h"] += n * 12
    img_w = pp_rec_padded_width(sizes.cpu().numpy())
    img_w[widths.cpu().numpy() <= 0] = 0
    ids_all = torch.full((n, t_max), -1, dtype=torch.int32, device=dev)
    lens_all = torch.zeros((n,), dtype=torch.int32, device=dev)
    conf_all = torch.zeros((n,), dtype=torch.float32, device=dev)
    for w in np.unique(img_w[img_w > 0]):
        idx = np.nonzero(img_w == w)[0]
        if len(idx) == n:
            g, gw = crops[:, :, : int(w)].contiguous(), widths
            idx_dev = None
        else:
            idx_dev = _h2d(torch.from_numpy(idx), dev)
            g, gw = crops[idx_dev, :, : int(w)], widths[idx_dev]
        ids, maxp = rec.rec_forward_u8(g, gw)
        out, ln, conf = post.ctc_collapse(ids, maxp)
        t = int(out.shape[1])
        if idx_dev is None:
            ids_all[:, :t], lens_all, conf_all = out, ln, conf
        else:
            ids_all[idx_dev, :t] = out
            lens_all[idx_dev] = ln
            conf_all[idx_dev] = conf
    return ids_all, lens_all, conf_all


class PPOcrRecPreProcessor:
    """Counterpart of PPOcrRecPreProcessor (ocr_rec_pp/processor_ocr_rec_pp.py:25-135): same call, same
    return value (a list of ``{'image', 'indices', 'batch_beg_img_no'}``), with ``image`` a CUDA fp32 [B,3,48,W] tensor.
    ``cv2.resize`` stays on the host exactly as in the reference; the float conversion, HWC->CHW, /255, -0.5, /0.5 and the
    zero padding run in one kernel (``dv_pp_rec_normalise``), bit-exact."""

    def __init__(self, engine: Engine, rec_image_shape=(3, 48, 320), rec_batch_num: int = 6, limited_max_width: int = 1280,
                 limited_min_width: int = 16, return_u8: bool = False):
        self.engine = engine
        self.return_u8 = return_u8
        self.rec_image_shape = tuple(rec_image_shape)
        self.rec_batch_num = rec_batch_num
        self.limited_max_width, self.limited_min_width = limited_max_width, limited_min_width

    def __call__(self, inputs):
        import cv2

        if not isinstance(inputs, list):
            inputs = [inputs]
        imgs = []
        for item in inputs:
            img = _read_image(item)
            if img.ndim == 2:
                img = cv2.cvtColor(img, cv2.COLOR_GRAY2RGB)
            imgs.append(img)
        img_c, img_h, _ = self.rec_image_shape
        indices, plan = pp_rec_batch_plan([im.shape[:2] for im in imgs], self.rec_image_shape, self.rec_batch_num,
                                          self.limited_min_width, self.limited_max_width)
        dev = torch.device("cuda", self.engine.device)
        out = []
        for beg, img_w, widths in plan:
            assert imgs[indices[beg]].shape[2] == img_c
            stage = np.zeros((len(widths), img_h, img_w, 3), np.uint8)
            for k, rw in enumerate(widths):
                stage[k, :, :rw] = cv2.resize(imgs[indices[beg + k]], (rw, img_h))
            if self.return_u8:  # the recogniser fuses the normalisation: keep the resized uint8 crops + their widths
                out.append({"image_u8": stage, "widths": np.array(widths, np.int32), "indices": indices, "batch_beg_img_no": beg})
                continue
            image = self.engine.pp_rec_normalise(torch.from_numpy(stage).to(dev),
                                                 torch.tensor(widths, dtype=torch.int32, device=dev))
            out.append({"image": image, "indices": indices, "batch_beg_img_no": beg})
        return out


TRANSFER = {"h2d": 0, "d2h": 0}  # bytes moved by the predictors since the last reset (bench.py reports them per step)


def _d2h_async(t: torch.Tensor) -> torch.Tensor:
    """Device -> pinned host copy on the current stream without waiting for it (torch allocates the destination from its
    caching pinned-memory pool); the caller synchronises an event recorded afterwards before reading."""
    TRANSFER["d2h"] += t.numel() * t.element_size()
    return t.to("cpu", non_blocking=True)


def _d2h(t: torch.Tensor) -> np.ndarray:
    TRANSFER["d2h"] += t.numel() * t.element_size()
    return t.cpu().numpy()


def _h2d(t: torch.Tensor, dev: torch.device) -> torch.Tensor:
    """Host tensor -> device on the current stream (asynchronous when the source is pinned memory)."""
    TRANSFER["h2d"] += t.numel() * t.element_size()
    return t.to(dev, non_blocking=True)


def _to_device_u8(img, dev: torch.device) -> torch.Tensor:
    """uint8 ndarray (pageable or a view of pinned memory) or tensor -> cuda tensor, asynchronously when the source is pinned."""
    if isinstance(img, torch.Tensor):
        return img if img.is_cuda else _h2d(img, dev)
    return _h2d(torch.from_numpy(np.ascontiguousarray(img)), dev)


def _pack_u8(imgs: Sequence[np.ndarray]):
    """uint8 HWC images -> (one flat uint8 array, byte offsets int64 [n]).  Images that are consecutive views of one contiguous
    array (``list(batch)``) are used in place; anything else is concatenated once."""
    n = len(imgs)
    sizes = np.array([im.size for im in imgs], np.int64)
    offs = np.concatenate([[0], np.cumsum(sizes)[:-1]]).astype(np.int64)
    first = imgs[0]
    if n > 1 and first.base is not None and first.dtype == np.uint8 and all(
            im.base is first.base and im.flags["C_CONTIGUOUS"] and im.dtype == np.uint8 and
            im.__array_interface__["data"][0] == first.__array_interface__["data"][0] + int(offs[k]) for k, im in enumerate(imgs)):
        total = int(sizes.sum())
        flat = np.lib.stride_tricks.as_strided(first.reshape(-1), shape=(total,), strides=(1,), writeable=False)
        return flat, offs
    return np.concatenate([np.ascontiguousarray(im, dtype=np.uint8).reshape(-1) for im in imgs]), offs


def order_points_batch(quads: np.ndarray) -> np.ndarray:
    """``order_point`` for n quads at once ([n,8] or [n,4,2] -> float32 [n,4,2]): the same numpy operations applied along a
    batch axis (sum / divide for the centroid, arctan2, the default argsort per quad, the rotation when the first corner is
    right of the centroid), so each row is bit-identical to the per-quad function (checked in tests/test_system_cpu.py)."""
    arr = np.asarray(quads).reshape(-1, 4, 2)
    if arr.shape[0] == 0:
        return np.zeros((0, 4, 2), np.float32)
    centroid = np.sum(arr, 1) / 4
    theta = np.arctan2(arr[:, :, 1] - centroid[:, 1:2], arr[:, :, 0] - centroid[:, 0:1])
    idx = np.argsort(theta, axis=1)
    pts = np.take_along_axis(arr, idx[:, :, None], axis=1)
    rot = pts[:, 0, 0] > centroid[:, 0]
    pts[rot] = np.concatenate([pts[rot][:, 3:], pts[rot][:, :3]], axis=1)
    return pts.astype("float32")


def _load_state_dict(sd_or_path) -> Mapping[str, Any]:
    if isinstance(sd_or_path, (str, os.PathLike)):
        sd = torch.load(sd_or_path, map_location="cpu")
        if isinstance(sd, dict) and "state_dict" in sd:
            sd = sd["state_dict"]
        return {k[7:] if k.startswith("module.") else k: v for k, v in sd.items()}
    return sd_or_path


class BaseInferTask:
    """Counterpart of BaseInferTask (ocr_pdf/base_infer_task.py:30-125, 311-315)."""

    SUPPORTS_FP32X = False

    def __init__(self, task: str = "", model: str = "", predictor_type: str = "b200", device: Union[int, str] = 0,
                 output_dir: Optional[str] = None, debug: bool = False, lang: str = "en", precision: str = "fp16", **kwargs):
        if predictor_type != "b200":
            raise RuntimeError(f"predictor_type '{predictor_type}' is served by the reference itself; this package only "
                               "provides 'b200'")
        if precision not in ("fp16", "fp32x") or (precision == "fp32x" and not self.SUPPORTS_FP32X):
            raise RuntimeError(f"precision '{precision}' not supported by this predictor: 'fp16' = fp16 operands / fp32 accumulation "
                               "(the reference default); 'fp32x' = split-fp16 operand pairs, ~fp32 products (logits within 1e-3 of "
                               "the fp32 graph) where the model implements it")
        self.precision = precision
        self.task, self.model, self._predictor_type = task, model, predictor_type
        self.device = int(str(device).replace("cuda:", "")) if not isinstance(device, int) else device
        self.output_dir, self.debug, self.lang, self.kwargs = output_dir, debug, lang, kwargs
        self.predictor: Optional[Engine] = None
        self._construct_model(model)
        self._build_processor()

    # the five hooks of the reference (base_infer_task.py:95-125)
    def _construct_model(self, model):  # pragma: no cover - abstract
        raise NotImplementedError

    def _build_processor(self):
        pass

    def _preprocess(self, inputs):  # pragma: no cover - abstract
        raise NotImplementedError

    def _run_model(self, inputs, **kwargs):  # pragma: no cover - abstract
        raise NotImplementedError

    def _postprocess(self, inputs, **kwargs):  # pragma: no cover - abstract
        raise NotImplementedError

    def __call__(self, inputs, **kwargs):
        return self._postprocess(self._run_model(self._preprocess(inputs), **kwargs), **kwargs)


def dbnet_resize_shape(height: int, width: int, image_short_side: int = 736):
    """OCRDetectionPreprocessor.resize (db_net/processor_ocr_dbnet.py:47-57): the short side becomes image_short_side, the
    other one keeps the aspect ratio rounded UP to a multiple of 32.  Returns (new_height, new_width)."""
    if height < width:
        new_height = image_short_side
        new_width = int(math.ceil(new_height / height * width / 32) * 32)
    else:
        new_width = image_short_side
        new_height = int(math.ceil(new_width / width * height / 32) * 32)
    return new_height, new_width


class OcrDetectionTask(BaseInferTask):
    """OcrDetectionTask (ocr_pdf/ocr_detection_task.py:30-141).  The network is the in-tree DBNet-R18 (db_net/dbnet.py:715)
    for both models (the PP-OCR det ONNX graph is not part of the reference repository); ``model`` selects the
    pre / post-processing of the reference back-end:

    * ``"db_pp"`` -- PPOcrDetectionPreprocessor / PPOcrDetectionPostProcessor (db_pp/processor_ocr_db_pp.py): DetResizeForTest
      (max side 960, multiples of 32), ``(x / 255 - mean) / std`` with the ImageNet statistics, box_thresh 0.6, boxes re-ordered
      clockwise, clipped and size-filtered (filter_tag_det_res).  Running DBNet-R18 weights under it is a stand-in for the
      PP-OCRv4 detector, not something the reference does.
    * ``"db"`` -- OCRDetectionPreprocessor / OCRDetectionPostProcessor (db_net/processor_ocr_dbnet.py:31-127): short side 736 with
      the other side rounded up to a multiple of 32, ``(x - [123.68, 116.78, 103.94]) / 255`` on the BGR-flipped image, box
      score >= 0.3, the mini box truncated to int32 then ``round(x / width * dest)``, no re-ordering or size filter
      (db_net/ocr_detection_utils.py:168-205).  This is the pipeline a real DBNet-R18 checkpoint was trained for.

    Returns list[np.ndarray [n, 8]] like the reference (:135-141)."""

    SUPPORTS_FP32X = True
    MEAN = (0.485, 0.456, 0.406)
    STD = (0.229, 0.224, 0.225)
    DB_MEAN = (123.68, 116.78, 103.94)  # OCRDetectionPreprocessor.normalize (db_net/processor_ocr_dbnet.py:59-62)

    def __init__(self, task: str = "ocr_detection", model: str = "db_pp", backbone: str = "resnet18", thresh: float = 0.2,
                 state_dict=None, box_thresh: Optional[float] = None, unclip_ratio: float = 1.5, max_candidates: int = 1000,
                 limit_side_len: int = 960, limit_type: str = "max", image_short_side: int = 736, **kwargs):
        if model not in ("db", "db_pp"):
            raise RuntimeError(f"model {model} not support")  # ocr_detection_task.py:58
        if state_dict is None:
            raise RuntimeError("OcrDetectionTask(predictor_type='b200') needs state_dict= (a DBModel state_dict or a path)")
        # backbone="resnet18": the in-tree DBModel (db_net/dbnet.py:715); backbone="PPLCNetV3" (model="db_pp" only): the PP-OCRv4
        # mobile detector the reference's db_pp back-end downloads as ONNX (PPLCNetV3-0.75 + RSE-FPN + DBHead, pp_det_graph.py)
        if backbone not in ("resnet18", "PPLCNetV3") or (backbone == "PPLCNetV3" and model != "db_pp"):
            raise RuntimeError(f"backbone {backbone} not support for model {model}")
        self.backbone = backbone
        self.thresh, self.unclip_ratio, self.max_candidates = thresh, unclip_ratio, max_candidates
        self.box_thresh = box_thresh if box_thresh is not None else (0.3 if model == "db" else 0.6)
        self.limit_side_len, self.limit_type, self.image_short_side = limit_side_len, limit_type, image_short_side
        self._sd = _load_state_dict(state_dict)
        super().__init__(task=task, model=model, **kwargs)

    def _construct_model(self, model):
        if self.backbone == "PPLCNetV3":
            from . import pp_det_graph

            self.predictor = Engine("pp_det", pp_det_graph.pack_pp_det(self._sd, precise=self.precision == "fp32x"), device=self.device)
        else:
            self.predictor = Engine("dbnet_r18", weights.pack_dbnet_r18(self._sd, precise=self.precision == "fp32x"), device=self.device)
        self._sd = None

    def _norm(self):
        """(mean, std, scale) of the fused normalisation kernel: (x * scale - mean) / std in float32."""
        if self.model == "db":
            return self.DB_MEAN, (255.0, 255.0, 255.0), 1.0
        return self.MEAN, self.STD, 1.0 / 255.0

    def _resize_host(self, img: np.ndarray):
        import cv2

        if self.model == "db":
            new_h, new_w = dbnet_resize_shape(img.shape[0], img.shape[1], self.image_short_side)
            return cv2.resize(img, (new_w, new_h)), [new_h / float(img.shape[0]), new_w / float(img.shape[1])]
        return det_resize_for_test(img, self.limit_side_len, self.limit_type)

    def _resize_device(self, page: torch.Tensor):
        if self.model == "db":
            new_h, new_w = dbnet_resize_shape(int(page.shape[0]), int(page.shape[1]), self.image_short_side)
            res = page if (new_h, new_w) == tuple(page.shape[:2]) else self.predictor.resize_pages_u8(page.unsqueeze(0), new_w, new_h)[0]
            return res, [new_h / float(page.shape[0]), new_w / float(page.shape[1])]
        return det_resize_for_test_device(self.predictor, page, self.limit_side_len, self.limit_type)

    def _preprocess(self, inputs) -> Dict[str, Any]:
        """PPOcrDetectionPreprocessor.__call__ (db_pp/processor_ocr_db_pp.py:103-145) / OCRDetectionPreprocessor.__call__
        (db_net/processor_ocr_dbnet.py:64-99) up to the uint8 resize; the channel flip, the normalisation and the CHW layout
        run fused on the GPU (dv_dbnet_forward_u8).  cv2.resize acts per channel, so resizing before the flip gives the same
        pixels as the reference's flip-then-resize."""
        if isinstance(inputs, np.ndarray) and inputs.ndim == 4:
            items = list(inputs)
        else:
            items = inputs if isinstance(inputs, (list, tuple)) else [inputs]
        pages, shapes, orgs = [], [], []
        for it in items:
            if isinstance(it, torch.Tensor):  # a page that is already on the device (uint8 HWC): resized there, nothing goes up
                if not it.is_cuda or it.dtype != torch.uint8 or it.dim() != 3:
                    raise TypeError("tensor inputs must be uint8 HWC cuda tensors")
                src_h, src_w = int(it.shape[0]), int(it.shape[1])
                res, (ratio_h, ratio_w) = self._resize_device(it)
                pages.append(res.contiguous())
                shapes.append(np.array([src_h, src_w, ratio_h, ratio_w]))
                orgs.append(tuple(it.shape))
                continue
            img = _read_image(it)
            src_h, src_w = img.shape[:2]
            res, (ratio_h, ratio_w) = self._resize_host(img)
            pages.append(np.ascontiguousarray(res))
            shapes.append(np.array([src_h, src_w, ratio_h, ratio_w]))
            orgs.append(img.shape)
        return {"pages": pages, "shape_list": shapes, "org_shape": orgs, "inputs": inputs}

    def _run_model(self, inputs, prob_override=None, **kwargs):
        """Launches DBNet + the whole box post-process per group of equally sized pages and starts the device -> host copies;
        nothing here waits for the GPU except the overflow read-back of dv_db_boxes.  ``prob_override`` (bench / tests with
        seeded random weights, whose probability map is texture noise): callable(prob [n,1,h,w] cuda, page indices) ->
        the map the box stage should read instead; the network still runs."""
        dev = torch.device("cuda", self.device)
        groups: Dict[tuple, List[int]] = {}
        for i, p in enumerate(inputs["pages"]):
            groups.setdefault((int(p.shape[0]), int(p.shape[1])), []).append(i)
        mean, std, scale = self._norm()
        pending = []
        for (h, w), idx in groups.items():  # one launch sequence per distinct resized shape
            members = [inputs["pages"][i] for i in idx]
            if len(members) > 1 and all(isinstance(m, torch.Tensor) for m in members) and _is_batch_view(members):
                batch = _batch_of(members)
            elif any(isinstance(m, torch.Tensor) for m in members):  # pages resized on the device (mixed groups: the others go up one by one)
                batch = torch.stack([m if isinstance(m, torch.Tensor) else _to_device_u8(m, dev) for m in members])
            else:
                batch = torch.empty((len(members), h, w, 3), dtype=torch.uint8, device=dev)
                for j, m in enumerate(members):
                    TRANSFER["h2d"] += m.nbytes
                    batch[j].copy_(torch.from_numpy(m), non_blocking=True)
            prob = self.predictor.dbnet_forward_u8(batch, mean, std, scale, flip=True)
            if prob_override is not None:
                prob = prob_override(prob, idx)
            src = [(inputs["shape_list"][i][0], inputs["shape_list"][i][1]) for i in idx]
            boxes, counts, overflow = self.predictor.db_boxes(prob, src, self.thresh, self.box_thresh, self.unclip_ratio, self.max_candidates,
                                                              check_overflow=True, variant=self.model)
            if overflow:
                import warnings

                warnings.warn(f"OcrDetectionTask: {overflow} contour(s) with more than 2048 vertices were skipped by dv_db_boxes "
                              "(the reference's cv2 path has no such limit)", RuntimeWarning)
            pending.append((idx, boxes, counts, _d2h_async(boxes), _d2h_async(counts)))
        ev = torch.cuda.Event()
        ev.record()
        inputs["pending"], inputs["event"] = pending, ev
        return inputs

    def _postprocess(self, inputs, **kwargs) -> List[np.ndarray]:
        inputs["event"].synchronize()
        boxes_out: List[Optional[np.ndarray]] = [None] * len(inputs["pages"])
        for idx, _, _, boxes_h, counts_h in inputs["pending"]:
            boxes, counts = boxes_h.numpy(), counts_h.numpy()
            for j, i in enumerate(idx):
                b = boxes[j, : counts[j]].copy()
                boxes_out[i] = b.astype(np.int64) if self.model == "db" else b  # np.array(lists of python ints) there (:126)
        inputs["det_polygons"] = boxes_out
        return boxes_out


def _is_batch_view(members) -> bool:
    """True when the cuda tensors are consecutive slices of one contiguous batch (what ``list(batch)`` / ``batch[i]`` give)."""
    first = members[0]
    step = first.numel() * first.element_size()
    return all(m.is_contiguous() and m.shape == first.shape and m.data_ptr() == first.data_ptr() + k * step for k, m in enumerate(members)) \
        and first._base is not None and all(m._base is first._base for m in members)


def _batch_of(members) -> torch.Tensor:
    first = members[0]
    return torch.as_strided(first, (len(members),) + tuple(first.shape), (first.numel(),) + tuple(first.stride()))


class OcrRecognitionTask(BaseInferTask):
    """OcrRecognitionTask (ocr_pdf/ocr_recognition_task.py:28-136) for model="ConvNextViT", "PP-OCRv4" and "CRNN".
    Returns list[str] like the reference (:118-136).  `vocab` is the character list of the checkpoint's vocab file
    (label ids start at 2 because do_chunking is set, ocr_recognition/processor_ocr_recognition.py:137-145).

    model="CRNN" (crnn/modeling_crnn.py) follows the reference's configuration as written: OCRRecognitionConfig is built with
    its defaults for every model_scope recogniser (ocr_recognition_task.py:33-35), i.e. do_chunking=True and width 804, so the
    pre-processor hands the network three 32 x 300 chunks per crop and the task returns the FIRST chunk's string
    (`predict['preds'][0]`, :124-128).  do_chunking=False (constructor kwarg) is the full-width mode the model was published
    with: one 32 x 804 image per crop, label ids from 1."""

    SUPPORTS_FP32X = True

    def __init__(self, task: str = "ocr_recognition", model: str = "ConvNextViT", task_type: str = "general", state_dict=None,
                 vocab: Optional[Sequence[str]] = None, do_chunking: bool = True, **kwargs):
        if model not in ("ConvNextViT", "PP-OCRv4", "CRNN"):
            raise RuntimeError(f"model {model} not support")
        if model != "ConvNextViT" and kwargs.get("precision", "fp16") != "fp16" and model == "CRNN":
            raise RuntimeError("the CRNN recogniser runs in fp16 operand precision only")
        self.do_chunking = do_chunking if model == "CRNN" else True
        if state_dict is None:
            raise RuntimeError("OcrRecognitionTask(predictor_type='b200') needs state_dict= (the recogniser's state_dict or a path)")
        self._sd = _load_state_dict(state_dict)
        if model == "PP-OCRv4":
            # CTCLabelDecode (ocr_rec_pp/rec_postprocess.py:20-45, 163-165): character = ['blank'] + dict lines + [' '] (use_space_char)
            self.character = ["blank"] + list(vocab) + [" "] if vocab is not None else None
            self.label_mapping = None
        else:
            first = 2 if self.do_chunking else 1  # load_vocab (processor_ocr_recognition.py:137-145)
            self.label_mapping = {i + first: ch for i, ch in enumerate(vocab)} if vocab is not None else None
        super().__init__(task=task, model=model, **kwargs)
        self.post = Engine("post", device=self.device)
        if model == "PP-OCRv4":
            self._pp_pre = PPOcrRecPreProcessor(self.post, return_u8=True)
            if self.character is not None and len(self.character) != self.predictor.rec_num_classes:
                raise RuntimeError(f"PP-OCRv4: the dictionary gives {len(self.character)} classes, the network has {self.predictor.rec_num_classes}")

    def _construct_model(self, model):
        if model == "PP-OCRv4":
            from .pp_rec_graph import pack_pp_rec

            self.predictor = Engine("pp_rec", pack_pp_rec(self._sd, precise=self.precision == "fp32x"), device=self.device)
        elif model == "CRNN":
            self.predictor = Engine("crnn", weights.pack_crnn(self._sd), device=self.device)
        else:
            self.predictor = Engine("convnext_vit", weights.pack_convnext_vit(self._sd, precise=self.precision == "fp32x"), device=self.device)
        self._sd = None

    @staticmethod
    def pp_read(items) -> List[np.ndarray]:
        """Input decoding of PPOcrRecPreProcessor.__call__ (processor_ocr_rec_pp.py:69-98): path / PIL / ndarray, grey -> RGB."""
        imgs = []
        for it in items:
            img = _read_image(it)
            if img.ndim == 2:
                import cv2

                img = cv2.cvtColor(img, cv2.COLOR_GRAY2RGB)
            imgs.append(img)
        return imgs

    def pp_batches_device(self, imgs: Sequence[np.ndarray]):
        """a4 on the device: the batch plan (aspect sort, groups of six, padded width per group, resized width per crop) is the
        reference's host rule (``pp_rec_batch_plan``); the cv2.resize of every crop is dv_resize_linear_u8 (bit-exact against
        cv2) over ONE upload of the raw crops, written straight into the zero-padded uint8 batch the network's first kernel reads.
        Groups of equal padded width are merged.  Returns [(original indices, uint8 [n,48,W,3] cuda, widths int32 [n] cuda)]."""
        dev = torch.device("cuda", self.device)
        indices, plan = pp_rec_batch_plan([im.shape[:2] for im in imgs])
        flat, offs = _pack_u8(imgs)
        import warnings

        with warnings.catch_warnings():  # the in-place view of the caller's batch is read-only for torch: it is only uploaded
            warnings.simplefilter("ignore", UserWarning)
            src = _h2d(torch.from_numpy(flat), dev)
        by_w: Dict[int, List[Tuple[int, int]]] = {}
        for beg, img_w, widths in plan:
            by_w.setdefault(int(img_w), []).extend((int(indices[beg + k]), int(rw)) for k, rw in enumerate(widths))
        out = []
        for w, members in by_w.items():
            order = np.array([m[0] for m in members], np.int64)
            widths = np.array([m[1] for m in members], np.int32)
            sizes = np.array([[imgs[i].shape[1], imgs[i].shape[0]] for i in order], np.int32)
            crops = self.post.resize_linear_u8((src, _h2d(torch.from_numpy(offs[order]), dev), _h2d(torch.from_numpy(sizes), dev)), widths, 48, w)
            out.append((order, crops, _h2d(torch.from_numpy(widths), dev)))
        return out

    # ---- model="PP-OCRv4": PPOcrRecPreProcessor (a4) -> SVTR-LCNet (a5) -> CTC greedy decode (a6)
    def _pp_call(self, inputs):
        """OcrRecognitionTask.__call__ for the PaddleOCR models (ocr_recognition_task.py:62-136): aspect-sorted batches of six with
        their own padded width (a4: ``pp_batches_device``); batches of EQUAL width share one network launch (every op
        of the network is per crop, so the result of a crop depends only on its own pixels and the padded width); per-step
        arg-max / max probability come from the network's last kernel, dv_ctc_collapse decodes them on the device and the
        characters are looked up on the host.  Returns the text of EVERY crop in input order (the reference's post-processor
        keeps only one entry of a multi-crop list, processor_ocr_rec_pp.py:150-160 -- its orchestrator always passes one crop)."""
        items = inputs if isinstance(inputs, (list, tuple)) else [inputs]
        pending = []
        for order, crops, widths_dev in self.pp_batches_device(self.pp_read(items)):
            ids, maxp = self.predictor.rec_forward_u8(crops, widths_dev)
            out, ln, conf = self.post.ctc_collapse(ids, maxp)
            pending.append((order, _d2h_async(out), _d2h_async(ln), _d2h_async(conf)))
        torch.cuda.current_stream().synchronize()
        texts: List[Optional[str]] = [None] * len(items)
        self.last_confidences = [0.0] * len(items)
        for order, out, ln, conf in pending:
            out, ln, conf = out.numpy(), ln.numpy(), conf.numpy()
            for k, i in enumerate(order):
                seq = out[k, : ln[k]]
                texts[int(i)] = " ".join(str(int(v)) for v in seq) if self.character is None else "".join(self.character[int(v)] for v in seq)
                self.last_confidences[int(i)] = float(conf[k])
        return texts

    def _preprocess(self, inputs) -> Dict[str, Any]:
        items = inputs if isinstance(inputs, (list, tuple)) else [inputs]
        crops = []
        for it in items:
            img = _read_image(it)
            if img.ndim == 2:
                img = np.stack([img] * 3, -1)
            crops.append(keepratio_resize(img))
        wmax = max(c.shape[1] for c in crops)
        # zero padding of the reference's mask (:57-61), staged in pinned memory so that the upload does not block the host
        stage = torch.zeros((len(crops), 32, wmax, 3), dtype=torch.uint8, pin_memory=True)
        batch = stage.numpy()
        for i, c in enumerate(crops):
            batch[i, :, : c.shape[1]] = c
        return {"crops": stage, "inputs": inputs}

    CALL_CHUNK = 512  # crops per pre-process / launch round of __call__

    def __call__(self, inputs, **kwargs):
        """post(run(pre(x))) like the reference; a long list is walked in chunks of CALL_CHUNK crops so that the host
        pre-processing (cv2.resize, padding) of chunk k+1 overlaps the device work of chunk k (nothing waits for the GPU
        before the final _postprocess)."""
        if self.model == "PP-OCRv4":
            return self._pp_call(inputs)
        items = inputs if isinstance(inputs, (list, tuple)) else [inputs]
        if len(items) <= self.CALL_CHUNK:
            return self._postprocess(self._run_model(self._preprocess(inputs), **kwargs), **kwargs)
        runs = [self._run_model(self._preprocess(list(items[i:i + self.CALL_CHUNK])), **kwargs) for i in range(0, len(items), self.CALL_CHUNK)]
        out: List[str] = []
        for r in runs:
            out += self._postprocess(r, **kwargs)
        return out

    def _crnn_batch(self, crops: np.ndarray) -> np.ndarray:
        """The rest of OCRRecognitionPreprocessor.__call__ (processor_ocr_recognition.py:57-61, 103-113) for the padded uint8
        crops [n,32,w,3]: pad to 804, float / 255, three chunks at x0 = 0, 252, 504 (or the whole line), NCHW."""
        n, _, w, _ = crops.shape
        full = np.zeros((n, 32, 804, 3), np.uint8)
        full[:, :, :w] = crops
        x = full.astype(np.float32) / np.float32(255.0)
        if self.do_chunking:
            x = np.stack([x[:, :, (300 - 48) * i:(300 - 48) * i + 300] for i in range(3)], 1).reshape(n * 3, 32, 300, 3)
        return np.ascontiguousarray(x.transpose(0, 3, 1, 2))

    def _run_model(self, inputs, **kwargs):
        dev = torch.device("cuda", self.device)
        crops = inputs["crops"]
        if self.model == "CRNN":
            x = self._crnn_batch(crops.numpy() if isinstance(crops, torch.Tensor) else crops)
            ids = self.predictor.crnn_forward(_h2d(torch.from_numpy(x), dev))
            if self.do_chunking:
                ids = ids.view(-1, 3, ids.shape[1])[:, 0].contiguous()  # predict['preds'][0]: the first chunk of every crop
            out, ln, _ = self.post.ctc_collapse(ids)
            inputs["ids_dev"], inputs["len_dev"] = out, ln
            inputs["ids"], inputs["len"] = _d2h_async(out), _d2h_async(ln)
            inputs["event"] = torch.cuda.Event()
            inputs["event"].record()
            return inputs
        ids = self.predictor.convnextvit_forward_u8(_h2d(crops if isinstance(crops, torch.Tensor) else torch.from_numpy(crops), dev))
        out, ln, _ = self.post.ctc_collapse(ids)
        inputs["ids_dev"], inputs["len_dev"] = out, ln
        inputs["ids"], inputs["len"] = _d2h_async(out), _d2h_async(ln)
        inputs["event"] = torch.cuda.Event()
        inputs["event"].record()
        return inputs

    def recognize_page(self, page, positions) -> List[Optional[str]]:
        """All detected quads of one page: what the reference's orchestrator does per box (OcrCommonUtils.crop_image, then
        this task on the crop -- ocr_pdf/ocr_system_task.py:300-313), with the crops cut and resized on the device
        (``dv_crop_quads_for_rec``) so that only the page goes up and only token ids come back.  page: uint8 HWC ndarray or
        cuda tensor.  Returns one string per quad (None where the reference's crop would be empty)."""
        return self.recognize_pages([page], [positions])[0]

    def launch_pages(self, pages, positions_per_page) -> Dict[str, Any]:
        """The asynchronous half of ``recognize_pages``: uploads what is not resident, enqueues crop + recognise + collapse for
        the quads of ALL pages as one batch and starts the device -> host copies; returns the pending record for
        ``collect_pages``.  pages: uint8 cuda tensor [P,H,W,3], or a list of equally sized HWC ndarrays / cuda tensors."""
        dev = torch.device("cuda", self.device)
        if isinstance(pages, torch.Tensor) and pages.dim() == 4:
            batch = pages
        else:
            members = [_to_device_u8(p, dev) for p in pages]
            if len({tuple(m.shape) for m in members}) > 1:
                raise ValueError("recognize_pages: pages of one call must have the same shape (call once per shape)")
            batch = _batch_of(members) if len(members) > 1 and _is_batch_view(members) else torch.stack(members)
        n_per = [len(p) for p in positions_per_page]
        if len(n_per) != int(batch.shape[0]):
            raise ValueError("one list of quads per page")
        rec = {"n_per": n_per, "event": None}
        total = sum(n_per)
        if total == 0:
            return rec
        quads = np.concatenate([np.asarray(p, np.float32).reshape(-1, 4, 2) for p in positions_per_page if len(p)])
        page_idx = np.repeat(np.arange(len(n_per), dtype=np.int32), n_per)
        q_dev = _h2d(torch.from_numpy(quads), dev)
        pi_dev = _h2d(torch.from_numpy(page_idx), dev)
        # geometry, homography, warp and resize all on the device; width 0 = skipped quad
        if self.model == "PP-OCRv4":  # crop_image + resize_norm_img (48 high, the crop's own padded width), then a5 + a6 per width group
            crops, widths, sizes, _ = self.post.crop_quads_for_rec(batch, q_dev, pi_dev, dst_h=48, dst_w_pad=1280, width_rule=1)
            tok, ln, conf = pp_rec_launch_groups(self.predictor, self.post, crops, widths, sizes)
            rec["conf"] = _d2h_async(conf)
        else:
            crops, widths, _, _ = self.post.crop_quads_for_rec(batch, q_dev, pi_dev)
            ids = self.predictor.convnextvit_forward_u8(crops)
            tok, ln, _ = self.post.ctc_collapse(ids)
        rec.update(ids_dev=tok, len_dev=ln, widths_dev=widths, ids=_d2h_async(tok), len=_d2h_async(ln), widths=_d2h_async(widths))
        rec["event"] = torch.cuda.Event()
        rec["event"].record()
        return rec

    def collect_pages(self, rec) -> List[List[Optional[str]]]:
        """Waits for ``launch_pages``' copies and maps ids to text: one list of strings per page (None = empty crop)."""
        out: List[List[Optional[str]]] = []
        if rec["event"] is None:
            return [[] for _ in rec["n_per"]]
        rec["event"].synchronize()
        if self.model == "PP-OCRv4":
            ids, lens = rec["ids"].numpy(), rec["len"].numpy()
            texts = self._ids_to_texts(ids, lens)
            self.last_confidences = rec["conf"].numpy().tolist()
        else:
            texts = self._postprocess({"ids": rec["ids"].numpy(), "len": rec["len"].numpy()})
        widths = rec["widths"].numpy()
        o = 0
        for n in rec["n_per"]:
            out.append([t if w > 0 else None for t, w in zip(texts[o:o + n], widths[o:o + n])])
            o += n
        return out

    def _ids_to_texts(self, ids: np.ndarray, lens: np.ndarray) -> List[str]:
        """CTCLabelDecode's character mapping (ocr_rec_pp/rec_postprocess.py:84-110) for collapsed ids [B, T] with lengths [B].  A
        vocabulary of single characters (every PP-OCR dictionary) is mapped with ONE table look-up for the whole batch: the
        [B, T] array of 1-character strings is re-viewed as B strings of T characters and cut to its length (1280 crops: 4.8 ->
        0.4 ms of host time on the e2e path); anything else takes the per-token join."""
        if self.character is None:
            return [" ".join(str(int(v)) for v in row[:k]) for row, k in zip(ids, lens)]
        tab = getattr(self, "_char_tab", None)
        if tab is None:
            # entry 0 is CTCLabelDecode's 'blank' token (a word, never emitted: the collapse removes it, and it pads the rows): it
            # gets a one-character stand-in that no dictionary entry uses, and a row that shows it inside its length falls back
            single = all(isinstance(c, str) and len(c) == 1 for c in self.character[1:])
            self._blank_mark = next(c for c in map(chr, range(0xE000, 0xF8FF)) if c not in set(self.character))
            first = self.character[0] if len(self.character[0]) == 1 else self._blank_mark
            tab = self._char_tab = np.array([first] + list(self.character[1:]), dtype="<U1") if single else False
        t = ids.shape[1] if ids.ndim == 2 else 0
        if tab is not False and t > 0 and ids.size and int(ids.min()) >= -1 and int(ids.max()) < len(tab):
            # -1 pads the rows beyond their length (pp_rec_launch_groups): looked up as the blank stand-in, cut off below
            rows = np.ascontiguousarray(tab[np.maximum(ids, 0)]).view(f"<U{t}").reshape(-1)
            mark = self._blank_mark if len(self.character[0]) != 1 else None
            # a row made of NUL-free characters keeps its full width; lengths cut the padding off
            return [str(r)[:k] if len(r) >= k and (mark is None or mark not in str(r)[:k]) else "".join(self.character[int(v)] for v in row[:k])
                    for r, k, row in zip(rows, lens.tolist(), ids)]
        return ["".join(self.character[int(v)] for v in row[:k]) for row, k in zip(ids, lens)]

    def recognize_pages(self, pages, positions_per_page) -> List[List[Optional[str]]]:
        """``recognize_page`` for a batch of equally sized pages: every quad of every page goes through ONE crop launch and one
        recogniser pass (the page index of each quad rides along), instead of one pass per page."""
        return self.collect_pages(self.launch_pages(pages, positions_per_page))

    def _postprocess(self, inputs, **kwargs) -> List[str]:
        if inputs.get("event") is not None:
            inputs["event"].synchronize()
        ids = inputs["ids"].numpy() if isinstance(inputs["ids"], torch.Tensor) else inputs["ids"]
        lens = inputs["len"].numpy() if isinstance(inputs["len"], torch.Tensor) else inputs["len"]
        res = []
        for row, n in zip(ids, lens):
            seq = row[:n]
            if self.label_mapping is None:
                res.append(" ".join(str(int(v)) for v in seq))
                continue
            try:
                res.append("".join(self.label_mapping[int(v)] for v in seq))
            except KeyError:
                # the reference's post-processor raises KeyError here (label id 1 / an id outside the vocab file) and its
                # orchestrator catches the exception PER CROP and keeps "" (ocr_system_task.py:309-320): same outcome, per row
                res.append("")
        return res


def lore_affine(center, scale, out_w: int, out_h: int, inv: bool = False) -> np.ndarray:
    """get_affine_transform(center, scale, rot=0, output_size=(out_w, out_h), inv) of the reference
    (lore/lineless_table_process.py:403-438): the centre-anchored similarity between the source image and the network
    frame, from three point pairs through cv2.getAffineTransform (float64 2x3)."""
    import cv2

    c = np.asarray(center, np.float32)
    src_w = np.float32(scale)
    src = np.zeros((3, 2), np.float32)
    dst = np.zeros((3, 2), np.float32)
    src[0] = c
    src[1] = c + np.array([0.0, src_w * np.float32(-0.5)], np.float32)
    dst[0] = [out_w * 0.5, out_h * 0.5]
    dst[1] = np.array([out_w * 0.5, out_h * 0.5], np.float32) + np.array([0, out_w * -0.5], np.float32)
    d = src[0] - src[1]
    src[2] = src[1] + np.array([-d[1], d[0]], np.float32)
    d = dst[0] - dst[1]
    dst[2] = dst[1] + np.array([-d[1], d[0]], np.float32)
    return cv2.getAffineTransform(dst, src) if inv else cv2.getAffineTransform(src, dst)


def lore_affine_upper_left(center, scale, out_w: int, out_h: int, inv: bool = False) -> np.ndarray:
    """get_affine_transform_upper_left(center, scale, rot=0, output_size=(out_w, out_h), inv) of the reference
    (lore/lineless_table_process.py:441-468), the `wireless` configuration's frame: `center` maps to the output's origin and
    the point one `scale` along an axis maps one out_w along it (the axis is chosen by center[0] < center[1]; the
    pre-processor passes center = (0, 0), :74-79, so it is the y axis and the map is the scaling out_w / scale about the
    upper-left corner)."""
    import cv2

    c = np.asarray(center, np.float32)
    sc = np.float32(scale)
    src = np.zeros((3, 2), np.float32)
    dst = np.zeros((3, 2), np.float32)
    src[0] = c
    if c[0] < c[1]:
        src[1] = [sc, c[1]]
        dst[1] = [out_w, 0]
    else:
        src[1] = [c[0], sc]
        dst[1] = [0, out_w]
    d = src[0] - src[1]
    src[2] = src[1] + np.array([-d[1], d[0]], np.float32)
    d = dst[0] - dst[1]
    dst[2] = dst[1] + np.array([-d[1], d[0]], np.float32)
    return cv2.getAffineTransform(dst, src) if inv else cv2.getAffineTransform(src, dst)


def _lore_frame(height: int, width: int, upper_left: bool):
    """(centre, scale, affine function) of TableLorePreProcessor.process (lore/processer_lore.py:73-83)."""
    s = max(height, width) * 1.0
    if upper_left:
        return np.array([0, 0], dtype=np.float32), s, lore_affine_upper_left
    return np.array([width / 2.0, height / 2.0], dtype=np.float32), s, lore_affine


def lore_preprocess(img: np.ndarray, resolution=(1024, 1024), upper_left: bool = False):
    """TableLorePreProcessor.process (lore/processer_lore.py:66-109) up to the uint8 warp: centre-anchored similarity
    warp (scale = resolution / max(h, w); anchored at the upper-left corner for the `wireless` configuration) with
    cv2.warpAffine (bilinear, zero border).  The normalisation and the CHW
    layout run fused on the GPU.  Returns (uint8 [H,W,3], meta int64 [7] = update_meta :112-130)."""
    import cv2

    height, width = img.shape[:2]
    inp_h, inp_w = resolution
    c, s, affine = _lore_frame(height, width, upper_left)
    trans = affine(c, s, inp_w, inp_h)
    warped = cv2.warpAffine(np.ascontiguousarray(img), trans, (inp_w, inp_h), flags=cv2.INTER_LINEAR)
    meta = np.array([c[0], c[1], s, inp_h, inp_w, inp_h // 4, inp_w // 4]).astype(np.int64)  # .long(): cx, cy truncated
    return warped, meta


def invert_affine(m) -> np.ndarray:
    """The inversion cv2.warpAffine applies to its 2x3 matrix before mapping destination pixels (imgwarp.cpp, the branch
    without WARP_INVERSE_MAP), in the same double operations."""
    m = np.asarray(m, np.float64).copy().ravel()
    d = m[0] * m[4] - m[1] * m[3]
    d = 1.0 / d if d != 0 else 0.0
    a11, a22 = m[4] * d, m[0] * d
    m[0] = a11
    m[1] *= -d
    m[3] *= -d
    m[4] = a22
    b1 = -m[0] * m[2] - m[1] * m[5]
    b2 = -m[3] * m[2] - m[4] * m[5]
    m[2], m[5] = b1, b2
    return m.reshape(2, 3)


def lore_preprocess_device(engine: Engine, img, resolution=(1024, 1024), upper_left: bool = False):
    """lore_preprocess with the warp on the device (``dv_warp_affine_u8``, bit-exact against cv2.warpAffine): img is a uint8 HWC
    ndarray or cuda tensor; only the raw image goes up.  Returns (uint8 [H,W,3] cuda tensor, meta int64 [7])."""
    dev = torch.device("cuda", engine.device)
    t = img if isinstance(img, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(img)).to(dev)
    height, width = int(t.shape[0]), int(t.shape[1])
    inp_h, inp_w = resolution
    c, s, affine = _lore_frame(height, width, upper_left)
    warped = engine.warp_affine_u8(t, invert_affine(affine(c, s, inp_w, inp_h)), inp_w, inp_h)
    meta = np.array([c[0], c[1], s, inp_h, inp_w, inp_h // 4, inp_w // 4]).astype(np.int64)
    return warped, meta


def table_crop_rect(bbox, height: int, width: int):
    """The slice OcrCommonUtils.crop_image_by_box takes (utils/ocr/ocr_common_utils.py:279-280, diff=0):
    img[round(y1):round(y2), round(x1):round(x2)] with Python's round (half to even) and numpy's slice clamping, as
    (x0, y0, crop_w, crop_h).  Raises ValueError for an empty slice (the reference fails in cv2.imwrite on it)."""
    x1, y1, x2, y2 = (float(v) for v in bbox)
    ys, ye, _ = slice(round(y1), round(y2)).indices(height)
    xs, xe, _ = slice(round(x1), round(x2)).indices(width)
    if ye <= ys or xe <= xs:
        raise ValueError(f"empty table crop for bbox {list(bbox)} on a {height}x{width} page")
    return xs, ys, xe - xs, ye - ys


class OcrTableStructureTask(BaseInferTask):
    """OcrTableStructureTask (ocr_pdf/ocr_table_structure_task.py:50-271) for model="Lore", task_type="wtw"
    (DLA-34 + DCNv2 detector, wiz_rev corner snapping, two 4-layer logical-location transformers), task_type="ptn" (the same
    detector at 512 x 512, no corner snapping, 3-layer transformers fed with 2-D position embeddings) or task_type="wireless"
    (the ResNet-18 key-point detector of lore/lore_detector.py at 768 x 768 in the upper-left-anchored frame, 4-layer
    transformers with 2-D position embeddings) and model="CenterNet"
    (DLA-34 + plain IDA-up, vertex grouping; returns list[dict{polygons [n,8]}] like OCRTableCenterNetPostProcessor).
    Returns list[dict{polygons [n,8] float32 source pixels, logi [n,4] integer-valued, inputs}] like the reference's
    TableLorePostProcessor (lore/processer_lore.py:163-188).  `state_dict` = (detector, processor) state_dicts or paths
    (the reference's model_best.pth / processor_best.pth, lore/modeling_lore.py:88-101)."""

    K, MK = 3000, 5000  # process_detect_output (lore/lineless_table_process.py:593)
    SUPPORTS_FP32X = True  # Lore detector: split-fp16 operand pairs (csrc/lore_net.cu); the processor always runs split

    def __init__(self, task: str = "ocr_table_structure", model: str = "Lore", task_type: str = "wtw", state_dict=None,
                 table_structure_merge: bool = False, max_cells_per_image: int = 3000, **kwargs):
        if model not in ("Lore", "CenterNet"):
            raise RuntimeError(f"model {model} not support")
        if model == "Lore" and task_type not in ("wtw", "ptn", "wireless"):
            raise RuntimeError(f"task_type {task_type} not support (the b200 predictor implements 'wtw', 'ptn' and 'wireless')")
        if model == "Lore" and (state_dict is None or len(state_dict) != 2):
            raise RuntimeError("OcrTableStructureTask(model='Lore', predictor_type='b200') needs state_dict=(detector, processor)")
        if model == "CenterNet" and state_dict is None:
            raise RuntimeError("OcrTableStructureTask(model='CenterNet', predictor_type='b200') needs state_dict= (a DLASeg state_dict or a path)")
        if model == "CenterNet" and kwargs.get("precision", "fp16") != "fp16":
            raise RuntimeError("the CenterNet detector runs in fp16 operand precision only")
        self.task_type, self.table_structure_merge = task_type, table_structure_merge
        self.upper_left = False
        if task_type == "ptn":  # LoreConfig ptn (configuration_lore.py:101-116): 512 x 512, 3 + 3 layers, 2-D position embeddings
            self.resolution, self.vis_thresh, self.wiz_rev, self.wiz_2dpe = (512, 512), 0.35, False, True
        elif task_type == "wireless" and model == "Lore":  # configuration_lore.py:73-85: ResNet-18, 768 x 768, upper-left frame
            self.resolution, self.vis_thresh, self.wiz_rev, self.wiz_2dpe, self.upper_left = (768, 768), 0.2, False, True, True
        else:  # LoreConfig wtw (configuration_lore.py:86-100)
            self.resolution, self.vis_thresh, self.wiz_rev, self.wiz_2dpe = (1024, 1024), 0.2, True, False
        # capacity of the cell-feature / processor buffers per image; the decode keeps at most K = 3000 cells per image
        # (process_detect_output), so the default can never overflow.  A smaller cap saves workspace; exceeding it raises.
        self.max_cells_per_image = max_cells_per_image
        self._sd = (_load_state_dict(state_dict[0]), _load_state_dict(state_dict[1])) if model == "Lore" else _load_state_dict(state_dict)
        super().__init__(task=task, model=model, **kwargs)

    def _construct_model(self, model):
        self.post = Engine("post", device=self.device)
        if model == "CenterNet":
            self.predictor = Engine("centernet_dla34", weights.pack_centernet_dla34(self._sd), device=self.device)
        else:
            if self.upper_left:  # LoreModel.__init__ picks the detector by the backbone name (lore/modeling_lore.py:79-89)
                self.predictor = Engine("lore_resnet18", weights.pack_lore_resnet18(self._sd[0], precise=self.precision == "fp32x"), device=self.device)
            else:
                self.predictor = Engine("lore_dla34", weights.pack_lore_dla34(self._sd[0], precise=self.precision == "fp32x"), device=self.device)
            self.processor = Engine("lore_processor", weights.pack_lore_processor(self._sd[1]), device=self.device)
        self._sd = None

    def _preprocess(self, inputs, **kwargs) -> Dict[str, Any]:
        """TableLorePreProcessor.process (lore/processer_lore.py:66-109): the raw image goes up and the centre-anchored
        cv2.warpAffine runs on the device (dv_warp_affine_u8, bit-exact against cv2 -- tests/test_gpu_crop.py); the
        normalisation is fused into the network's first kernel.  host_warp=True (constructor kwarg) keeps the cv2 call on
        the host as the reference does."""
        items = inputs if isinstance(inputs, (list, tuple)) else [inputs]
        dev = torch.device("cuda", self.device)
        images, metas, cs = [], [], []
        for it in items:
            img = _read_image(it)
            if img.ndim == 2:
                img = np.stack([img] * 3, -1)
            if not isinstance(it, np.ndarray):
                img = img[:, :, ::-1]  # path / PIL inputs reach the network as BGR, ndarrays unchanged (processer_lore.py:51-60, 146)
            if self.kwargs.get("host_warp", False):
                warped, meta = lore_preprocess(img, self.resolution, self.upper_left)
                warped = _to_device_u8(warped, dev)
            else:
                warped, meta = lore_preprocess_device(self.post, _to_device_u8(img, dev), self.resolution, self.upper_left)
            images.append(warped)
            metas.append(meta)
            h, w = img.shape[:2]
            cs.append((np.array([w / 2.0, h / 2.0], dtype=np.float32), max(h, w) * 1.0))
        return {"images": torch.stack(images), "meta": np.stack(metas), "cs": cs, "inputs": list(items)}

    def _images_on_device(self, images) -> torch.Tensor:
        """The warped uint8 batch: uploaded when the host pre-process made it, used in place when recognize_tables did."""
        if isinstance(images, torch.Tensor) and images.is_cuda:
            return images
        return _h2d(images if isinstance(images, torch.Tensor) else torch.from_numpy(images), torch.device("cuda", self.device))

    def recognize_tables(self, pages, layout_tables) -> List[list]:
        """The table loop of the reference's orchestrator (ocr_pdf/ocr_system_task.py:184-198) for pages that are already on
        the device: per layout table, OcrCommonUtils.crop_image_by_box + this task on the crop -- here all crops of all pages
        are cut and warped into the network frame by ONE launch (``dv_crop_tables_for_tsr``: slice + cv2.warpAffine, bit-exact
        on the page's pixels; the reference's lossy JPEG write / re-read of the crop is not reproduced), so only 68 bytes per
        table go up and only the decoded cells come back.
        pages: uint8 HWC ndarray / cuda tensor [H,W,3] or a batch [P,H,W,3], in the channel order the crops should reach the
        network in (the reference re-reads its JPEG as RGB and flips it to BGR, processer_lore.py:51-60, 146).
        layout_tables: dicts with "bbox" = [x1,y1,x2,y2] (and "page" when pages is a batch), e.g. the table rows of
        OcrLayoutTask's result.  Returns [[bbox, result], ...] in input order like ``outputs`` there (:189-198), result being
        what ``self(crop)[0]`` returns for that crop (its "inputs" is the bbox)."""
        return self.collect_tables(self.launch_tables(pages, layout_tables))

    def launch_tables(self, pages, layout_tables) -> Optional[Dict[str, Any]]:
        """The asynchronous half of ``recognize_tables`` (crop + warp + network + decode + processor enqueued, copies started)."""
        dev = torch.device("cuda", self.device)
        t = pages if isinstance(pages, torch.Tensor) else _h2d(torch.from_numpy(np.ascontiguousarray(pages)), dev)
        if t.dim() == 3:
            t = t.unsqueeze(0)
        n_pages, height, width = int(t.shape[0]), int(t.shape[1]), int(t.shape[2])
        if len(layout_tables) == 0:
            return None
        inp_h, inp_w = self.resolution
        rects, minv, metas, cs = [], [], [], []
        for tb in layout_tables:
            pg = int(tb.get("page", 0))
            if not 0 <= pg < n_pages:
                raise ValueError(f"table page {pg} outside the batch of {n_pages} pages")
            x0, y0, cw, ch = table_crop_rect(tb["bbox"], height, width)
            c, sc, affine = _lore_frame(ch, cw, self.upper_left)
            rects.append([pg, x0, y0, cw, ch])
            minv.append(invert_affine(affine(c, sc, inp_w, inp_h)))
            metas.append(np.array([c[0], c[1], sc, inp_h, inp_w, inp_h // 4, inp_w // 4]).astype(np.int64))
            cs.append((c, sc))
        warped = self.post.crop_tables_for_tsr(t, np.array(rects, np.int32), np.stack(minv), inp_w, inp_h)
        items = [tb["bbox"] for tb in layout_tables]
        return self._run_model({"images": warped, "meta": np.stack(metas), "cs": cs, "inputs": items})

    def collect_tables(self, pending) -> List[list]:
        if pending is None:
            return []
        res = self._postprocess(pending)
        return [[bbox, r] for bbox, r in zip(pending["inputs"], res)]

    def _run_centernet(self, inputs):
        """OCRTableCenterNetPreProcessor keeps the float centre / scale in its meta (center_net/processer_centernet.py:108-139);
        the pixels are the same warp as Lore's."""
        inv = np.stack([lore_affine(c, s, self.resolution[1] // 4, self.resolution[0] // 4, inv=True) for c, s in inputs["cs"]])
        maps = self.predictor.lore_detect_forward_u8(self._images_on_device(inputs["images"]))
        polygons, counts = self.post.centernet_decode(maps, None, None, None, inv)
        inputs["dev"] = {"polygons": polygons, "counts": counts}
        inputs["host"] = {"polygons": _d2h_async(polygons), "counts": _d2h_async(counts)}
        inputs["event"] = torch.cuda.Event()
        inputs["event"].record()
        return inputs

    def _run_model(self, inputs, **kwargs):
        """Enqueues detector -> decode -> cell features -> processor on the current stream (no host round trip between them:
        the row counts stay on the device) and starts the device -> host copies of the fixed-capacity result buffers."""
        if self.model == "CenterNet":
            return self._run_centernet(inputs)
        n = len(inputs["images"])
        metas = inputs["meta"]
        # ctdet_4ps_post_process / _upper_left (lineless_table_process.py:493-530) on the integer meta (update_meta's .long())
        affine = lore_affine_upper_left if self.upper_left else lore_affine
        inv = np.stack([affine([np.float32(m[0]), np.float32(m[1])], np.float32(m[2]), int(m[6]), int(m[5]), inv=True) for m in metas])
        maps = self.predictor.lore_detect_forward_u8(self._images_on_device(inputs["images"]))
        dec = self.post.lore_decode(maps, None, None, None, inv, K=self.K, MK=self.MK, wiz_rev=self.wiz_rev, vis_thresh=self.vis_thresh)
        cap = n * min(self.max_cells_per_image, self.K)
        feat, offsets = self.predictor.lore_cell_features(dec, max_rows=cap)
        if self.wiz_2dpe:  # LoreModel.forward passes dets=slct_dets_feat (modeling_lore.py:155-159)
            self.processor.lore_add_position_embeddings(feat, dec, offsets)
        _, stacked = self.processor.lore_process_forward(feat, offsets)
        inputs["dev"] = {"polygons": dec["polygons"], "counts": dec["counts"], "offsets": offsets, "logi": stacked}
        inputs["cap"] = cap
        inputs["host"] = {"counts": _d2h_async(dec["counts"]), "offsets": _d2h_async(offsets)}
        inputs["event"] = torch.cuda.Event()
        inputs["event"].record()
        return inputs

    def _collect_lore(self, inputs):
        """Second half of the Lore run: the counts are on the host now, so only the used part of the polygon / coordinate
        buffers is copied back."""
        inputs["event"].synchronize()
        counts = inputs["host"]["counts"].numpy()
        offs = inputs["host"]["offsets"].numpy()
        n = len(counts)
        if n and int(counts.sum()) > inputs["cap"]:
            raise DocVisionError(f"lore: {int(counts.sum())} cells exceed the buffer capacity {inputs['cap']} "
                                 f"(max_cells_per_image={self.max_cells_per_image}); construct the task with a larger cap")
        kmax = int(counts.max()) if n else 0
        polygons = _d2h(inputs["dev"]["polygons"][:, :max(kmax, 1)])
        logits = _d2h(inputs["dev"]["logi"][: int(offs[-1])])
        return [{"pred_boxes": polygons[i, : counts[i]], "logits": logits[offs[i]: offs[i + 1]]} for i in range(n)]

    def _postprocess(self, inputs, **kwargs) -> List[Dict[str, Any]]:
        if self.model == "CenterNet":  # {"polygons": np.array(box_list), ...inputs} (processer_centernet.py:198-203)
            inputs["event"].synchronize()
            polygons, counts = inputs["host"]["polygons"].numpy(), inputs["host"]["counts"].numpy()
            return [{"polygons": polygons[i, : counts[i]].copy(), "inputs": item} for i, item in enumerate(inputs["inputs"])]
        out = []
        for item, res in zip(inputs["inputs"], self._collect_lore(inputs)):
            boxes, logi = res["pred_boxes"], res["logits"]
            if len(boxes) == 0:  # LoreModel.forward (lore/modeling_lore.py:171-173)
                boxes, logi = np.zeros((1, 8), np.float32), np.zeros((1, 4), np.float32)
            f = np.floor(logi)
            logi = np.where(logi - f > 0.5, f + 1, f)  # process_logic_output (lineless_table_process.py:658-663)
            out.append({"polygons": boxes, "logi": logi.astype(np.float32), "inputs": item})
        return out


class OcrLayoutTask(BaseInferTask):
    """OcrLayoutTask (ocr_pdf/ocr_layout_task.py:27-157) for model="picodet": LCNet-x1.0 + CSP-PAN + PicoHead on the engine,
    anchor decode + per-class NMS on the GPU.  Returns list[list[dict{bbox float64[4], label, score, category_id}]] like the
    reference (picodet/processor_picodet.py:288-297).  `state_dict` = (backbone, neck, head) state_dicts of the reference
    modules (the hub ships only the ONNX export of this architecture)."""

    LABELS = {  # PicodetConfig.label_config (picodet/configuration_picodet.py:83-108)
        "ch": ["text", "title", "figure", "figure_caption", "table", "table_caption", "header", "footer", "reference", "equation"],
        "en": ["text", "title", "list", "table", "figure"],
        "table": ["table"],
    }
    IMG_H, IMG_W = 800, 608

    def __init__(self, task: str = "ocr_layout", model: str = "picodet", task_type: str = "en", score_threshold: float = 0.5,
                 nms_threshold: float = 0.5, state_dict=None, nms_top_k: int = 1000, keep_top_k: int = 100, **kwargs):
        if model != "picodet":
            raise RuntimeError(f"model {model} not support")
        if task_type not in ("ch", "en", "table"):
            task_type = "en"  # ocr_layout_task.py:36-37
        if state_dict is None or len(state_dict) != 3:
            raise RuntimeError("OcrLayoutTask(predictor_type='b200') needs state_dict=(backbone, neck, head)")
        self.task_type, self.score_threshold, self.nms_threshold = task_type, score_threshold, nms_threshold
        self.nms_top_k, self.keep_top_k = nms_top_k, keep_top_k
        self.id2label = dict(enumerate(self.LABELS[task_type]))
        self._sd = tuple(_load_state_dict(s) for s in state_dict)
        super().__init__(task=task, model=model, **kwargs)

    def _construct_model(self, model):
        from .picodet_graph import pack_picodet

        self.predictor = Engine("picodet", pack_picodet(*self._sd, num_classes=len(self.id2label)), device=self.device)
        self.post = Engine("post", device=self.device)
        self._sd = None

    def _preprocess(self, inputs, **kwargs) -> Dict[str, Any]:
        """OCRPicodetPreProcessor.__call__ (picodet/processor_picodet.py:72-113) up to the uint8 resize (the reference
        flips the channels first and resizes afterwards; cv2.resize acts per channel, so resizing first and flipping on the
        GPU gives the same pixels)."""
        import cv2

        items = inputs if isinstance(inputs, (list, tuple)) else [inputs]
        imgs, org, sf = [], [], []
        if len(items) > 1 and all(isinstance(it, torch.Tensor) and it.is_cuda and it.dtype == torch.uint8 and it.dim() == 3 for it in items) \
                and _is_batch_view(list(items)):  # the pages of one resident batch: ONE resize launch for all of them
            h, w = int(items[0].shape[0]), int(items[0].shape[1])
            images = self.post.resize_pages_u8(_batch_of(list(items)), self.IMG_W, self.IMG_H)
            return {"images": images, "org_shape": [(h, w)] * len(items), "scale_factor": [(float(self.IMG_H) / h, float(self.IMG_W) / w)] * len(items),
                    "inputs": list(items)}
        for it in items:
            if isinstance(it, torch.Tensor):  # a page that is already on the device (uint8 HWC): resized there
                if not it.is_cuda or it.dtype != torch.uint8 or it.dim() != 3:
                    raise TypeError("tensor inputs must be uint8 HWC cuda tensors")
                h, w = int(it.shape[0]), int(it.shape[1])
                imgs.append(self.post.resize_pages_u8(it.unsqueeze(0), self.IMG_W, self.IMG_H)[0])
            else:
                img = _read_image(it)
                h, w = img.shape[:2]
                imgs.append(cv2.resize(np.ascontiguousarray(img), (self.IMG_W, self.IMG_H)))
            org.append((h, w))
            sf.append((float(self.IMG_H) / h, float(self.IMG_W) / w))
        if any(isinstance(m, torch.Tensor) for m in imgs):
            dev = torch.device("cuda", self.device)
            images = torch.stack([m if isinstance(m, torch.Tensor) else torch.from_numpy(m).to(dev) for m in imgs])
        else:
            images = np.stack(imgs)
        return {"images": images, "org_shape": org, "scale_factor": sf, "inputs": list(items)}

    def _run_model(self, inputs, **kwargs):
        dev = torch.device("cuda", self.device)
        images = inputs["images"]
        if not isinstance(images, torch.Tensor):
            images = _h2d(torch.from_numpy(images), dev)
        scores, dfl = self.predictor.picodet_forward_u8(images, flip=True)
        boxes, counts = self.post.picodet_decode(scores, dfl, inputs["org_shape"], inputs["scale_factor"], (self.IMG_H, self.IMG_W),
                                                 score_threshold=self.score_threshold, nms_threshold=self.nms_threshold,
                                                 nms_top_k=self.nms_top_k, keep_top_k=self.keep_top_k)
        inputs["boxes"], inputs["counts"] = _d2h_async(boxes), _d2h_async(counts)
        inputs["event"] = torch.cuda.Event()
        inputs["event"].record()
        return inputs

    def _postprocess(self, inputs, **kwargs) -> List[List[Dict[str, Any]]]:
        inputs["event"].synchronize()
        out = []
        for rows, n in zip(inputs["boxes"].numpy(), inputs["counts"].numpy()):
            out.append([{"bbox": r[2:].copy(), "label": self.id2label[int(r[0])], "score": r[1], "category_id": int(r[0])} for r in rows[:n]])
        return out


class ClsImagePulcTask(BaseInferTask):
    """ClsImagePulcTask (ocr_pdf/cls_image_pulc_task.py:23-100): the PULC PP-LCNet classifiers the reference's
    orchestrator runs before the cascade for image inputs (ocr_system_task.py:395-439) -- task_type "text_image_orientation",
    "textline_orientation", "language_classification" (Topk post-process) or "table_attribute" (thresholds).  The host keeps the
    reference's image processor steps (PIL bilinear resize through transformers.image_transforms, rescale 1/255, ImageNet
    normalise, CHW: cls/image_processing_pplcnet.py:327-455) and its post-processors (:109-193); the network runs on the engine
    (dv_cls_forward, graph program pplcnet_graph.py).  Returns what the reference returns: one result dict per input, the bare
    dict for a single input (:94-96).  `state_dict` = the PPLCNet module's state_dict or a path."""

    CLASS_ID_MAP = {  # cls/image_processing_pplcnet.py:41-72
        "text_image_orientation": {0: "0", 1: "90", 2: "180", 3: "270"},
        "textline_orientation": {0: "0_degree", 1: "180_degree"},
        "language_classification": {0: "arabic", 1: "chinese_cht", 2: "cyrillic", 3: "devanagari", 4: "japan", 5: "ka", 6: "korean",
                                    7: "ta", 8: "te", 9: "latin"},
    }
    IMAGE_SIZE = {"text_image_orientation": (224, 224), "textline_orientation": (80, 160), "language_classification": (80, 160)}
    TOPK = {"text_image_orientation": 2, "textline_orientation": 1, "language_classification": 2}

    def __init__(self, task: str = "cls_image", model: str = "PPLCNet", task_type: str = "text_image_orientation", state_dict=None, **kwargs):
        from .pplcnet_graph import TASK_CLASSES

        if model != "PPLCNet" or task_type not in TASK_CLASSES:
            raise RuntimeError(f"model {model} / task_type {task_type} not support")
        if state_dict is None:
            raise RuntimeError("ClsImagePulcTask(predictor_type='b200') needs state_dict= (a PPLCNet state_dict or a path)")
        self.task_type = task_type
        self._sd = _load_state_dict(state_dict)
        super().__init__(task=task, model=model, **kwargs)

    def _construct_model(self, model):
        from .pplcnet_graph import TASK_CLASSES, TASK_STRIDES, pack_pplcnet

        if int(np.asarray(self._sd["fc.weight"]).shape[0]) != TASK_CLASSES[self.task_type]:
            raise RuntimeError(f"task_type {self.task_type} expects {TASK_CLASSES[self.task_type]} classes")
        self.predictor = Engine("pplcnet_cls", pack_pplcnet(self._sd, TASK_STRIDES[self.task_type]), device=self.device)
        self._sd = None

    def _preprocess(self, inputs, **kwargs):
        from transformers.image_transforms import normalize, rescale, resize, to_channel_dimension_format
        from transformers.image_utils import ChannelDimension, PILImageResampling, to_numpy_array
        from transformers.utils import IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD

        items = inputs if isinstance(inputs, list) else [inputs]
        size = self.IMAGE_SIZE.get(self.task_type, (224, 224))
        batch = []
        for it in items:
            if isinstance(it, str):  # preprocess() reads paths with cv2 and flips to RGB (:389-397)
                import cv2

                img = cv2.imread(it)[:, :, ::-1]
            else:
                img = it
            img = to_numpy_array(img)
            img = resize(img, size=size, resample=PILImageResampling.BILINEAR)
            img = normalize(rescale(img, 1 / 255), IMAGENET_DEFAULT_MEAN, IMAGENET_DEFAULT_STD)
            batch.append(np.asarray(to_channel_dimension_format(img, ChannelDimension.FIRST), np.float32))
        return {"pixel_values": np.stack(batch), "inputs": items}

    def _run_model(self, inputs, **kwargs):
        dev = torch.device("cuda", self.device)
        logits = self.predictor.cls_forward(_h2d(torch.from_numpy(inputs["pixel_values"]), dev))
        inputs["logits"] = _d2h(logits)
        return inputs

    def _postprocess(self, inputs, **kwargs):
        logits = torch.from_numpy(inputs["logits"])
        results = []
        if self.task_type == "table_attribute":  # TableAttribute.__call__ (:125-153): thresholds 0.5 on the network's raw outputs
            names = (("Scanned", "Photo"), ("Little", "Numerous"), ("Black-and-White", "Multicolor"), ("Clear", "Blurry"),
                     ("Without-Obstacles", "With-Obstacles"), ("Horizontal", "Tilted"))
            for res in logits.tolist():
                results.append({"attributes": [a if v > 0.5 else b for v, (a, b) in zip(res, names)],
                                "output": (np.array(res) > np.array([0.5] * 6)).astype(np.int8).tolist()})
        else:  # Topk.__call__ (:162-192)
            probs_all = torch.nn.functional.softmax(logits, dim=-1).numpy()
            id_map = self.CLASS_ID_MAP[self.task_type]
            for probs in probs_all:
                index = probs.argsort(axis=0)[-self.TOPK[self.task_type]:][::-1].astype("int32")
                results.append({"class_ids": [i.item() for i in index],
                                "scores": np.around([probs[i].item() for i in index], decimals=5).tolist(),
                                "label_names": [id_map[i.item()] for i in index]})
        return results[0] if len(results) == 1 else results
