"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement of the PP-OCR recogniser pre-process, SURVEY.md a4:
PPOcrRecPreProcessor.__call__ / resize_norm_img (ocr_rec_pp/processor_ocr_rec_pp.py:44-135) with the defaults of
PPOcrRecognitionConfig (configuration_ocr_recognition_pp.py:45-48: rec_image_shape "3, 48, 320", limited widths 16 .. 1280,
rec_batch_num 6).  Pinned against the reference class itself by oracle/gen_golden_pp_rec_pre.py ->
tests/golden/pp_rec_pre.npz (tests/test_pp_rec_pre_cpu.py).
"""
from __future__ import annotations

import math

import cv2
import numpy as np

IMG_C, IMG_H, IMG_W = 3, 48, 320
MIN_W, MAX_W, BATCH = 16, 1280, 6


def batch_plan(shapes):
    """shapes: [(h, w)] -> (indices, [(beg, img_w, [resized_w per crop of the batch])]).  processor_ocr_rec_pp.py:100-121 for
    the order and the batches, :44-59 for the widths."""
    ratios = np.array([w / float(h) for h, w in shapes])
    indices = np.argsort(ratios)  # same call as the reference (ties: numpy's introsort order)
    plan = []
    for beg in range(0, len(shapes), BATCH):
        ids = indices[beg:beg + BATCH]
        max_ratio = 0
        for i in ids:
            h, w = shapes[i]
            max_ratio = max(max_ratio, w * 1.0 / h)
        max_ratio = max(max_ratio, IMG_W / IMG_H)
        img_w = max(min(int(IMG_H * max_ratio), MAX_W), MIN_W)
        widths = []
        for i in ids:
            h, w = shapes[i]
            rw = max(math.ceil(IMG_H * (w / float(h))), MIN_W)
            widths.append(img_w if rw > img_w else int(rw))
        plan.append((beg, img_w, widths))
    return indices, plan


def preprocess(crops):
    """list of uint8 HWC crops -> list of {'image' fp32 [B,3,48,W], 'indices', 'batch_beg_img_no'} as the reference returns."""
    indices, plan = batch_plan([c.shape[:2] for c in crops])
    out = []
    for beg, img_w, widths in plan:
        batch = np.zeros((len(widths), IMG_C, IMG_H, img_w), np.float32)
        for k, rw in enumerate(widths):
            r = cv2.resize(crops[indices[beg + k]], (rw, IMG_H)).astype("float32")
            r = r.transpose((2, 0, 1)) / 255
            r -= 0.5
            r /= 0.5
            batch[k, :, :, :rw] = r
        out.append({"image": batch, "indices": indices, "batch_beg_img_no": beg})
    return out
