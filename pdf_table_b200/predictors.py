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
from typing import Any, Dict, List, Mapping, Optional, Sequence, Union

import numpy as np
import torch

from . import weights
from .engine import Engine

__all__ = ["BaseInferTask", "OcrDetectionTask", "OcrRecognitionTask", "det_resize_for_test", "keepratio_resize"]


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
    if (resize_h, resize_w) != (h, w):
        img = cv2.resize(img, (int(resize_w), int(resize_h)))
    return img, [resize_h / float(h), resize_w / float(w)]


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


def _load_state_dict(sd_or_path) -> Mapping[str, Any]:
    if isinstance(sd_or_path, (str, os.PathLike)):
        sd = torch.load(sd_or_path, map_location="cpu")
        if isinstance(sd, dict) and "state_dict" in sd:
            sd = sd["state_dict"]
        return {k[7:] if k.startswith("module.") else k: v for k, v in sd.items()}
    return sd_or_path


class BaseInferTask:
    """Counterpart of BaseInferTask (ocr_pdf/base_infer_task.py:30-125, 311-315)."""

    def __init__(self, task: str = "", model: str = "", predictor_type: str = "b200", device: Union[int, str] = 0,
                 output_dir: Optional[str] = None, debug: bool = False, lang: str = "en", precision: str = "fp16", **kwargs):
        if predictor_type != "b200":
            raise RuntimeError(f"predictor_type '{predictor_type}' is served by the reference itself; this package only "
                               "provides 'b200'")
        if precision != "fp16":
            raise RuntimeError("the b200 predictor computes with fp16 operands / fp32 accumulation (the reference default)")
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


class OcrDetectionTask(BaseInferTask):
    """OcrDetectionTask (ocr_pdf/ocr_detection_task.py:30-141).  model="db" / "db_pp" select the pre/post-processing
    constants of the two reference back-ends; the network is the in-tree DBNet-R18 (db_net/dbnet.py:715) either way
    (the PP-OCR det ONNX graph is not part of the reference repository).
    Returns list[np.ndarray [n, 8]] like the reference (:135-141)."""

    MEAN = (0.485, 0.456, 0.406)
    STD = (0.229, 0.224, 0.225)

    def __init__(self, task: str = "ocr_detection", model: str = "db_pp", backbone: str = "resnet18", thresh: float = 0.2,
                 state_dict=None, box_thresh: float = 0.6, unclip_ratio: float = 1.5, max_candidates: int = 1000,
                 limit_side_len: int = 960, limit_type: str = "max", **kwargs):
        if model not in ("db", "db_pp"):
            raise RuntimeError(f"model {model} not support")  # ocr_detection_task.py:58
        if state_dict is None:
            raise RuntimeError("OcrDetectionTask(predictor_type='b200') needs state_dict= (a DBModel state_dict or a path)")
        self.thresh, self.box_thresh, self.unclip_ratio, self.max_candidates = thresh, box_thresh, unclip_ratio, max_candidates
        self.limit_side_len, self.limit_type = limit_side_len, limit_type
        self._sd = _load_state_dict(state_dict)
        super().__init__(task=task, model=model, **kwargs)

    def _construct_model(self, model):
        self.predictor = Engine("dbnet_r18", weights.pack_dbnet_r18(self._sd), device=self.device)
        self._sd = None

    def _preprocess(self, inputs) -> Dict[str, Any]:
        """PPOcrDetectionPreprocessor.__call__ (db_pp/processor_ocr_db_pp.py:103-145) up to the uint8 resize; the
        channel flip, NormalizeImage and ToCHWImage run fused on the GPU (dv_dbnet_forward_u8)."""
        items = inputs if isinstance(inputs, (list, tuple)) else [inputs]
        pages, shapes, orgs = [], [], []
        for it in items:
            img = _read_image(it)
            src_h, src_w = img.shape[:2]
            res, (ratio_h, ratio_w) = det_resize_for_test(img, self.limit_side_len, self.limit_type)
            pages.append(np.ascontiguousarray(res))
            shapes.append(np.array([src_h, src_w, ratio_h, ratio_w]))
            orgs.append(img.shape)
        return {"pages": pages, "shape_list": shapes, "org_shape": orgs, "inputs": inputs}

    def _run_model(self, inputs, **kwargs):
        dev = torch.device("cuda", self.device)
        groups: Dict[tuple, List[int]] = {}
        for i, p in enumerate(inputs["pages"]):
            groups.setdefault(p.shape[:2], []).append(i)
        boxes_out: List[Optional[np.ndarray]] = [None] * len(inputs["pages"])
        for (h, w), idx in groups.items():  # one launch sequence per distinct resized shape
            batch = torch.from_numpy(np.stack([inputs["pages"][i] for i in idx])).to(dev, non_blocking=True)
            prob = self.predictor.dbnet_forward_u8(batch, self.MEAN, self.STD, 1.0 / 255.0, flip=True)
            src = [(inputs["shape_list"][i][0], inputs["shape_list"][i][1]) for i in idx]
            boxes, counts = self.predictor.db_boxes(prob, src, self.thresh, self.box_thresh, self.unclip_ratio, self.max_candidates)
            boxes, counts = boxes.cpu().numpy(), counts.cpu().numpy()
            for j, i in enumerate(idx):
                boxes_out[i] = boxes[j, : counts[j]].copy()
        inputs["det_polygons"] = boxes_out
        return inputs

    def _postprocess(self, inputs, **kwargs) -> List[np.ndarray]:
        return inputs["det_polygons"]


class OcrRecognitionTask(BaseInferTask):
    """OcrRecognitionTask (ocr_pdf/ocr_recognition_task.py:28-136) for model="ConvNextViT".
    Returns list[str] like the reference (:118-136).  `vocab` is the character list of the checkpoint's vocab file
    (label ids start at 2 because do_chunking is set, ocr_recognition/processor_ocr_recognition.py:137-145)."""

    def __init__(self, task: str = "ocr_recognition", model: str = "ConvNextViT", task_type: str = "general", state_dict=None,
                 vocab: Optional[Sequence[str]] = None, **kwargs):
        if model != "ConvNextViT":
            raise RuntimeError(f"model {model} not support")
        if state_dict is None:
            raise RuntimeError("OcrRecognitionTask(predictor_type='b200') needs state_dict= (a ConvNextViT state_dict or a path)")
        self._sd = _load_state_dict(state_dict)
        self.label_mapping = {i + 2: ch for i, ch in enumerate(vocab)} if vocab is not None else None
        super().__init__(task=task, model=model, **kwargs)
        self.post = Engine("post", device=self.device)

    def _construct_model(self, model):
        self.predictor = Engine("convnext_vit", weights.pack_convnext_vit(self._sd), device=self.device)
        self._sd = None

    def _preprocess(self, inputs) -> Dict[str, Any]:
        items = inputs if isinstance(inputs, (list, tuple)) else [inputs]
        crops = []
        for it in items:
            img = _read_image(it)
            if img.ndim == 2:
                img = np.stack([img] * 3, -1)
            crops.append(keepratio_resize(img))
        wmax = max(c.shape[1] for c in crops)
        batch = np.zeros((len(crops), 32, wmax, 3), np.uint8)  # zero padding of the reference's mask (:57-61)
        for i, c in enumerate(crops):
            batch[i, :, : c.shape[1]] = c
        return {"crops": batch, "inputs": inputs}

    def _run_model(self, inputs, **kwargs):
        dev = torch.device("cuda", self.device)
        ids = self.predictor.convnextvit_forward_u8(torch.from_numpy(inputs["crops"]).to(dev, non_blocking=True))
        out, ln, _ = self.post.ctc_collapse(ids)
        inputs["ids"], inputs["len"] = out.cpu().numpy(), ln.cpu().numpy()
        return inputs

    def _postprocess(self, inputs, **kwargs) -> List[str]:
        res = []
        for row, n in zip(inputs["ids"], inputs["len"]):
            seq = row[:n]
            if self.label_mapping is None:
                res.append(" ".join(str(int(v)) for v in seq))
            else:
                res.append("".join(self.label_mapping[int(v)] for v in seq))  # KeyError on id 1, as the reference
        return res
