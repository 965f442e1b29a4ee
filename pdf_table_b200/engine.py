"""Python handle over the C ABI.  torch is used only for device memory and streams (plumbing); every
computation below is a kernel in libdocvision.so.  There is no fallback: without the library or without
a B200 these calls raise DocVisionError."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import DocVisionError, check


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _require_cuda(t: torch.Tensor, dtype, name: str) -> torch.Tensor:
    if not t.is_cuda:
        raise DocVisionError(f"{name} must be a CUDA tensor (no CPU path exists)")
    if t.dtype != dtype:
        raise TypeError(f"{name} must be {dtype}, got {t.dtype}")
    return t.contiguous()


class _StreamBoundLib:
    """The ctypes library as seen by one Engine: every dv_* call first points the handle at torch's CURRENT stream of the
    handle's device (dv_set_stream when it changed), so kernels are ordered with the torch copies / allocations issued
    around them even when the caller switches streams with ``torch.cuda.stream(s)`` after constructing the engine."""

    __slots__ = ("_lib", "_engine", "_cache")

    def __init__(self, lib, engine):
        self._lib, self._engine, self._cache = lib, engine, {}

    def __getattr__(self, name):
        fn = self._cache.get(name)
        if fn is None:
            raw = getattr(self._lib, name)
            eng = self._engine

            def fn(*args, _raw=raw, _eng=eng):
                _eng._bind_stream()
                return _raw(*args)

            self._cache[name] = fn
        return fn


def _params_to(a: np.ndarray, dev) -> torch.Tensor:
    """A small host parameter array -> device WITHOUT a stream synchronisation: `tensor.to(dev)` of pageable memory waits for the
    whole current stream after the copy, which parked the host behind the previous batch's device work once two batches were in
    flight (12.8 ms per step in OcrSystemTask.predict_stream); the non-blocking form returns once CUDA has staged the bytes."""
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=True)


class Engine:
    """One handle per (model kind, device).  Not thread-safe; distinct handles are independent."""

    def __init__(self, kind: str = "post", blob: Optional[bytes] = None, device: int = 0):
        self._raw = _lib.load()
        self._h = C.c_void_p()
        self.kind = kind
        self.device = device
        self._stream = None
        buf = (C.c_char * len(blob)).from_buffer_copy(blob) if blob else None
        rc = self._raw.dv_create(kind.encode(), buf, len(blob) if blob else 0, device, C.byref(self._h))
        check(rc, None, f"dv_create({kind})")
        self._lib = _StreamBoundLib(self._raw, self)
        self.use_current_stream()

    def _bind_stream(self):
        s = torch.cuda.current_stream(self.device).cuda_stream
        if s != self._stream and self._h.value:
            check(self._raw.dv_set_stream(self._h, C.c_void_p(s)), self._h, "dv_set_stream")
            self._stream = s

    # ------------------------------------------------------------------ lifetime / streams
    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._raw.dv_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def use_current_stream(self):
        """Bind the handle to torch's current stream of its device (also done implicitly before every call)."""
        self._stream = None
        self._bind_stream()

    def sync(self):
        check(self._lib.dv_sync(self._h), self._h, "dv_sync")

    @property
    def launch_count(self) -> int:
        return int(self._lib.dv_launch_count(self._h))

    @property
    def model_flops(self) -> float:
        return float(self._lib.dv_model_flops(self._h))

    def profile_begin(self):
        """Start bracketing every launch with CUDA events (see dv_profile_begin)."""
        check(self._lib.dv_profile_begin(self._h), self._h, "dv_profile_begin")

    def profile_report(self):
        """Stop profiling; list of {kernel, layer, ms, flops, bytes} per launch."""
        import json

        cap = 1 << 20
        while True:
            buf = C.create_string_buffer(cap)
            n = int(self._lib.dv_profile_report(self._h, buf, cap))
            if n < 0:
                check(n, self._h, "dv_profile_report")
            if n < cap:
                return json.loads(buf.value.decode())
            cap = n + 1024  # records are kept until a report fits

    # ------------------------------------------------------------------ networks
    def dbnet_forward(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """fp32 NCHW [N,3,H,W] (cuda) -> probability map fp32 [N,1,H,W]."""
        x = _require_cuda(x, torch.float32, "x")
        n, c, h, w = x.shape
        if c != 3:
            raise ValueError("dbnet expects 3 channels")
        if out is None:
            out = torch.empty((n, 1, h, w), dtype=torch.float32, device=x.device)
        check(self._lib.dv_dbnet_forward(self._h, _ptr(x), n, h, w, _ptr(out)), self._h, "dv_dbnet_forward")
        return out

    def dbnet_forward_u8(self, pages: torch.Tensor, mean, std, scale: float = 1.0 / 255.0, flip: bool = True,
                         out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """uint8 HWC pages [N,H,W,3] (cuda) -> probability map; fuses the reference's normalisation."""
        pages = _require_cuda(pages, torch.uint8, "pages")
        n, h, w, c = pages.shape
        if c != 3:
            raise ValueError("pages must be [N,H,W,3]")
        if out is None:
            out = torch.empty((n, 1, h, w), dtype=torch.float32, device=pages.device)
        m3 = (C.c_float * 3)(*[float(v) for v in mean])
        s3 = (C.c_float * 3)(*[float(v) for v in std])
        check(self._lib.dv_dbnet_forward_u8(self._h, _ptr(pages), n, h, w, m3, s3, float(scale), int(flip), _ptr(out)),
              self._h, "dv_dbnet_forward_u8")
        return out

    def convnextvit_forward(self, chunks: torch.Tensor, return_logits: bool = False, return_max: bool = False):
        """fp32 [3n,3,32,300] (cuda, in [0,1]) -> per-token arg-max ids int32 [n,201] (+ logits fp32 [n,201,L],
        + max logit [n,201])."""
        chunks = _require_cuda(chunks, torch.float32, "chunks")
        b3, c, hh, ww = chunks.shape
        if (c, hh, ww) != (3, 32, 300) or b3 % 3:
            raise ValueError("chunks must be [3n,3,32,300]")
        n = b3 // 3
        dev = chunks.device
        labels = int(self._lib.dv_convnextvit_labels(self._h))
        ids = torch.empty((n, 201), dtype=torch.int32, device=dev)
        logits = torch.empty((n, 201, labels), dtype=torch.float32, device=dev) if return_logits else None
        mx = torch.empty((n, 201), dtype=torch.float32, device=dev) if return_max else None
        check(self._lib.dv_convnextvit_forward(self._h, _ptr(chunks), n, _ptr(logits), _ptr(ids), _ptr(mx)), self._h,
              "dv_convnextvit_forward")
        out = (ids,)
        if return_logits:
            out += (logits,)
        if return_max:
            out += (mx,)
        return out if len(out) > 1 else ids

    def convnextvit_forward_u8(self, crops: torch.Tensor, ids: Optional[torch.Tensor] = None, return_logits: bool = False):
        """uint8 [n,32,w,3] (cuda; height 32, right-padded with zeros to a common w <= 804) -> arg-max ids [n,201]."""
        crops = _require_cuda(crops, torch.uint8, "crops")
        n, hh, w, c = crops.shape
        if hh != 32 or c != 3 or not (0 < w <= 804):
            raise ValueError("crops must be [n,32,w<=804,3]")
        dev = crops.device
        if ids is None:
            ids = torch.empty((n, 201), dtype=torch.int32, device=dev)
        labels = int(self._lib.dv_convnextvit_labels(self._h))
        logits = torch.empty((n, 201, labels), dtype=torch.float32, device=dev) if return_logits else None
        check(self._lib.dv_convnextvit_forward_u8(self._h, _ptr(crops), n, w, _ptr(logits), _ptr(ids), None), self._h,
              "dv_convnextvit_forward_u8")
        return (ids, logits) if return_logits else ids

    def crnn_forward(self, x: torch.Tensor, return_logits: bool = False, return_max: bool = False):
        """CRNN (model kind "crnn"): fp32 [n,3,32,W] (cuda, in [0,1], W % 4 == 0) -> per-step arg-max ids int32 [n, W/4]
        (+ logits fp32 [n, W/4, L], + max logit [n, W/4])."""
        x = _require_cuda(x, torch.float32, "x")
        n, c, hh, ww = x.shape
        if c != 3 or hh != 32 or ww % 4:
            raise ValueError("x must be [n,3,32,W] with W % 4 == 0")
        t = ww // 4
        labels = int(self._lib.dv_crnn_labels(self._h))
        ids = torch.empty((n, t), dtype=torch.int32, device=x.device)
        logits = torch.empty((n, t, labels), dtype=torch.float32, device=x.device) if return_logits else None
        mx = torch.empty((n, t), dtype=torch.float32, device=x.device) if return_max else None
        check(self._lib.dv_crnn_forward(self._h, _ptr(x), n, hh, ww, _ptr(logits), _ptr(ids), _ptr(mx)), self._h, "dv_crnn_forward")
        out = (ids,)
        if return_logits:
            out += (logits,)
        if return_max:
            out += (mx,)
        return out if len(out) > 1 else ids

    def set_pass_crops(self, crops: int):
        check(self._lib.dv_convnextvit_set_pass_crops(self._h, int(crops)), self._h, "dv_convnextvit_set_pass_crops")

    def warp_perspective_u8(self, page: torch.Tensor, minv: np.ndarray, sizes: np.ndarray, return_packed: bool = False):
        """uint8 HWC page (cuda) + per-crop inverse homographies [n,3,3] float64 and sizes [n,2] = (w, h) (host) -> list of
        uint8 [h,w,3] cuda tensors (views of one packed buffer): cv2.warpPerspective(page, T, (w, h)) for each crop.
        return_packed: also return (packed buffer, device offsets int64 [n], device sizes int32 [n,2]) for resize_linear_u8."""
        page = _require_cuda(page, torch.uint8, "page")
        hh, ww, c = page.shape
        minv = np.ascontiguousarray(minv, dtype=np.float64).reshape(-1, 9)
        sizes = np.ascontiguousarray(sizes, dtype=np.int32).reshape(-1, 2)
        n = sizes.shape[0]
        if c != 3 or minv.shape[0] != n:
            raise ValueError("page must be [H,W,3]; one 3x3 matrix and one (w, h) per crop")
        if n == 0:
            return []
        if (sizes <= 0).any():
            raise ValueError("crop sizes must be positive")
        nbytes = sizes[:, 0].astype(np.int64) * sizes[:, 1] * 3
        offsets = np.concatenate([[0], np.cumsum(nbytes)[:-1]]).astype(np.int64)
        dev = page.device
        out = torch.empty((int(nbytes.sum()),), dtype=torch.uint8, device=dev)
        d_m = _params_to(minv, dev)
        d_s = _params_to(sizes, dev)
        d_o = _params_to(offsets, dev)
        check(self._lib.dv_warp_perspective_u8(self._h, _ptr(page), hh, ww, _ptr(d_m), _ptr(d_s), _ptr(d_o), n,
                                               int((nbytes // 3).max()), _ptr(out)), self._h, "dv_warp_perspective_u8")
        views = [out[int(o):int(o + b)].view(int(s[1]), int(s[0]), 3) for o, b, s in zip(offsets, nbytes, sizes)]
        return (views, (out, d_o, d_s)) if return_packed else views

    def resize_linear_u8(self, packed, dst_widths: np.ndarray, dst_h: int, dst_w_pad: int) -> torch.Tensor:
        """cv2.resize(crop_i, (dst_widths[i], dst_h)) (INTER_LINEAR) for the packed crops of warp_perspective_u8
        (``packed`` = its (buffer, offsets, sizes) triple) -> uint8 [n, dst_h, dst_w_pad, 3], zero beyond each width."""
        buf, d_o, d_s = packed
        n = d_s.shape[0]
        dst_widths = np.ascontiguousarray(dst_widths, dtype=np.int32).reshape(-1)
        if dst_widths.shape[0] != n or (dst_widths <= 0).any() or (dst_widths > dst_w_pad).any():
            raise ValueError("one positive width <= dst_w_pad per crop")
        out = torch.empty((n, dst_h, dst_w_pad, 3), dtype=torch.uint8, device=buf.device)
        d_w = _params_to(dst_widths, buf.device)
        check(self._lib.dv_resize_linear_u8(self._h, _ptr(buf), _ptr(d_o), _ptr(d_s), _ptr(d_w), n, dst_h, dst_w_pad, _ptr(out)),
              self._h, "dv_resize_linear_u8")
        return out

    def resize_pages_u8(self, pages: torch.Tensor, dst_w: int, dst_h: int) -> torch.Tensor:
        """uint8 [n,H,W,3] (cuda) -> uint8 [n,dst_h,dst_w,3]: cv2.resize(page, (dst_w, dst_h)) (INTER_LINEAR) for same-size pages,
        through dv_resize_linear_u8 (the pages are the packed sources: offset i*H*W*3, size (W, H) each)."""
        pages = _require_cuda(pages, torch.uint8, "pages")
        if pages.dim() != 4 or pages.shape[3] != 3:
            raise ValueError("pages must be [n,H,W,3]")
        n, hh, ww, _ = pages.shape
        if n == 0 or dst_w <= 0 or dst_h <= 0:
            raise ValueError("resize_pages_u8: empty batch / bad size")
        dev = pages.device
        offs = _params_to((np.arange(n, dtype=np.int64) * (hh * ww * 3)), dev)
        sizes = _params_to(np.array([[ww, hh]] * n, dtype=np.int32), dev)
        return self.resize_linear_u8((pages, offs, sizes), np.full((n,), dst_w, np.int32), dst_h, dst_w)

    def crop_quads_for_rec(self, pages: torch.Tensor, quads: torch.Tensor, page_idx: Optional[torch.Tensor] = None, dst_h: int = 32,
                           dst_w_pad: int = 804, width_rule: int = 0):
        """pages uint8 [P,H,W,3] (or [H,W,3]) + quads float32 [n,4,2] (+ page index int32 [n]), all cuda -> (crops uint8
        [n,dst_h,dst_w_pad,3], widths int32 [n] (0 = skipped), crop sizes int32 [n,2], inverse homographies float64 [n,3,3]),
        all on the device: crop_image + keepratio_resize of the reference for every quad, no host round trip."""
        pages = _require_cuda(pages, torch.uint8, "pages")
        if pages.dim() == 3:
            pages = pages.unsqueeze(0)
        quads = _require_cuda(quads, torch.float32, "quads").reshape(-1, 8)
        n = quads.shape[0]
        pp, hh, ww, c = pages.shape
        if c != 3:
            raise ValueError("pages must be [P,H,W,3]")
        if page_idx is not None:
            page_idx = _require_cuda(page_idx, torch.int32, "page_idx")
            if page_idx.numel() != n:
                raise ValueError("one page index per quad")
        dev = pages.device
        out = torch.empty((n, dst_h, dst_w_pad, 3), dtype=torch.uint8, device=dev)
        widths = torch.empty((n,), dtype=torch.int32, device=dev)
        minv = torch.empty((n, 3, 3), dtype=torch.float64, device=dev)
        sizes = torch.empty((n, 2), dtype=torch.int32, device=dev)
        check(self._lib.dv_crop_quads_for_rec(self._h, _ptr(pages), pp, hh, ww, _ptr(quads), _ptr(page_idx), n, dst_h, dst_w_pad,
                                              _ptr(out), _ptr(widths), _ptr(minv), _ptr(sizes), int(width_rule)), self._h, "dv_crop_quads_for_rec")
        return out, widths, sizes, minv

    def crop_boxes_for_rec(self, pages: torch.Tensor, boxes: torch.Tensor, counts: torch.Tensor, per_page: int, dst_h: int = 32,
                           dst_w_pad: int = 804, out: Optional[torch.Tensor] = None, ws=None, width_rule: int = 0):
        """crop_quads_for_rec on db_boxes' own outputs: boxes float32 [P,S,8] + counts int32 [P] (cuda) -> per_page crop slots per
        page, (crops uint8 [P*per_page,dst_h,dst_w_pad,3], widths int32 [P*per_page], sizes, minv).  ``out`` / ``ws`` = (widths,
        sizes, minv) may be passed in to reuse buffers (no allocation on the hot path)."""
        pages = _require_cuda(pages, torch.uint8, "pages")
        boxes = _require_cuda(boxes, torch.float32, "boxes")
        counts = _require_cuda(counts, torch.int32, "counts")
        pp, hh, ww, c = pages.shape
        if c != 3 or boxes.dim() != 3 or boxes.shape[0] != pp or boxes.shape[2] != 8 or counts.numel() != pp:
            raise ValueError("pages [P,H,W,3], boxes [P,S,8], counts [P]")
        n = pp * per_page
        dev = pages.device
        if out is None:
            out = torch.empty((n, dst_h, dst_w_pad, 3), dtype=torch.uint8, device=dev)
        if ws is None:
            ws = (torch.empty((n,), dtype=torch.int32, device=dev), torch.empty((n, 2), dtype=torch.int32, device=dev),
                  torch.empty((n, 3, 3), dtype=torch.float64, device=dev))
        widths, sizes, minv = ws
        check(self._lib.dv_crop_boxes_for_rec(self._h, _ptr(pages), pp, hh, ww, _ptr(boxes), _ptr(counts), boxes.shape[1], per_page, dst_h,
                                              dst_w_pad, _ptr(out), _ptr(widths), _ptr(minv), _ptr(sizes), int(width_rule)), self._h, "dv_crop_boxes_for_rec")
        return out, widths, sizes, minv

    def warp_affine_u8(self, img: torch.Tensor, m_inv: np.ndarray, out_w: int, out_h: int) -> torch.Tensor:
        """uint8 HWC image (cuda) + the inverted 2x3 matrix (host float64) -> uint8 [out_h,out_w,3]: cv2.warpAffine, bilinear."""
        img = _require_cuda(img, torch.uint8, "img")
        hh, ww, c = img.shape
        if c != 3:
            raise ValueError("img must be [H,W,3]")
        m = np.ascontiguousarray(m_inv, dtype=np.float64).reshape(6)
        out = torch.empty((out_h, out_w, 3), dtype=torch.uint8, device=img.device)
        check(self._lib.dv_warp_affine_u8(self._h, _ptr(img), hh, ww, m.ctypes.data_as(C.c_void_p), out_w, out_h, _ptr(out)), self._h,
              "dv_warp_affine_u8")
        return out

    def crop_tables_for_tsr(self, pages: torch.Tensor, rects: np.ndarray, m_inv: np.ndarray, out_w: int, out_h: int) -> torch.Tensor:
        """uint8 pages [P,H,W,3] (cuda), rects int32 [n,5] = page, x0, y0, w, h and inverted matrices float64 [n,2,3] (host; 68
        bytes per table go up) -> uint8 [n,out_h,out_w,3]: every table slice warped as cv2.warpAffine warps the cut-out crop."""
        pages = _require_cuda(pages, torch.uint8, "pages")
        if pages.dim() != 4 or pages.shape[3] != 3:
            raise ValueError("pages must be [P,H,W,3]")
        rects = np.ascontiguousarray(rects, dtype=np.int32).reshape(-1, 5)
        m = np.ascontiguousarray(m_inv, dtype=np.float64).reshape(-1, 6)
        n = rects.shape[0]
        if m.shape[0] != n:
            raise ValueError("one inverted matrix per rect")
        out = torch.empty((n, out_h, out_w, 3), dtype=torch.uint8, device=pages.device)
        if n == 0:
            return out
        rects_d = _params_to(rects, pages.device)
        m_d = _params_to(m, pages.device)
        check(self._lib.dv_crop_tables_for_tsr(self._h, _ptr(pages), pages.shape[0], pages.shape[1], pages.shape[2], _ptr(rects_d), _ptr(m_d), n,
                                               out_w, out_h, _ptr(out)), self._h, "dv_crop_tables_for_tsr")
        return out

    def pp_rec_normalise(self, crops: torch.Tensor, widths: torch.Tensor) -> torch.Tensor:
        """uint8 [B,H,W,3] resized crops (left-aligned, widths int32 [B]) -> fp32 [B,3,H,W]: (x/255 - 0.5)/0.5, zero padded."""
        crops = _require_cuda(crops, torch.uint8, "crops")
        widths = _require_cuda(widths, torch.int32, "widths")
        b, hh, ww, c = crops.shape
        if c != 3 or widths.numel() != b:
            raise ValueError("crops must be [B,H,W,3] with one width per crop")
        out = torch.empty((b, 3, hh, ww), dtype=torch.float32, device=crops.device)
        check(self._lib.dv_pp_rec_normalise(self._h, _ptr(crops), _ptr(widths), b, hh, ww, _ptr(out)), self._h,
              "dv_pp_rec_normalise")
        return out

    def ctc_collapse(self, ids: torch.Tensor, scores: Optional[torch.Tensor] = None, blank: int = 0):
        """[B,T] int32 per-step arg-max (+ optional [B,T] fp32 scores) -> (ids left-packed/-1 padded, len, conf)."""
        ids = _require_cuda(ids, torch.int32, "ids")
        if scores is not None:
            scores = _require_cuda(scores, torch.float32, "scores")
        b, t = ids.shape
        dev = ids.device
        out = torch.empty((b, t), dtype=torch.int32, device=dev)
        ln = torch.empty((b,), dtype=torch.int32, device=dev)
        conf = torch.empty((b,), dtype=torch.float32, device=dev)
        check(self._lib.dv_ctc_collapse(self._h, _ptr(ids), _ptr(scores), b, t, blank, _ptr(out), _ptr(ln), _ptr(conf)),
              self._h, "dv_ctc_collapse")
        return out, ln, conf

    def match_cells(self, text_boxes: torch.Tensor, cell_boxes: torch.Tensor) -> torch.Tensor:
        """float64 [T,4] text boxes and [C,4] table cells (cuda; x1, y1, x2, y2) -> int32 [T]: the cell of every text box
        (find_top1_mach_box of the reference's table export)."""
        text_boxes = _require_cuda(text_boxes, torch.float64, "text_boxes")
        cell_boxes = _require_cuda(cell_boxes, torch.float64, "cell_boxes")
        if text_boxes.dim() != 2 or text_boxes.shape[1] != 4 or cell_boxes.dim() != 2 or cell_boxes.shape[1] != 4 or cell_boxes.shape[0] == 0:
            raise ValueError("text_boxes [T,4], cell_boxes [C>0,4]")
        out = torch.empty((text_boxes.shape[0],), dtype=torch.int32, device=text_boxes.device)
        check(self._lib.dv_match_cells(self._h, _ptr(text_boxes), int(text_boxes.shape[0]), _ptr(cell_boxes), int(cell_boxes.shape[0]), _ptr(out)),
              self._h, "dv_match_cells")
        return out

    def debug_tensor(self, name: str) -> torch.Tensor:
        """Named intermediate activation of the last forward as fp32 NCHW (parity debugging)."""
        dims = (C.c_int * 4)()
        check(self._lib.dv_debug_get_tensor(self._h, name.encode(), None, dims), self._h, "dv_debug_get_tensor")
        out = torch.empty(tuple(dims), dtype=torch.float32, device=f"cuda:{self.device}")
        check(self._lib.dv_debug_get_tensor(self._h, name.encode(), _ptr(out), dims), self._h, "dv_debug_get_tensor")
        return out

    # ------------------------------------------------------------------ post-processing kernels
    def db_boxes(self, prob: torch.Tensor, src_hw, thresh: float = 0.2, box_thresh: float = 0.6, unclip_ratio: float = 1.5,
                 max_candidates: int = 1000, check_overflow: bool = False, variant: str = "db_pp"):
        """prob fp32 [N,1,H,W] (cuda); src_hw = [(src_h, src_w)] per page -> (boxes fp32 [N,max_candidates,8], counts
        int32 [N]) on the device.  Boxes follow the reference's order and corner convention.  variant "db_pp" =
        PPOcrDetectionPostProcessor (dv_db_boxes), "db" = the in-tree DBNet back-end's OCRDetectionPostProcessor
        (dv_db_boxes_dbnet: int32 truncation before scaling, no clockwise re-ordering / size filter)."""
        if variant not in ("db_pp", "db"):
            raise ValueError(f"db_boxes variant {variant!r}")
        prob = _require_cuda(prob, torch.float32, "prob")
        n, c, h, w = prob.shape
        if c != 1:
            raise ValueError("prob must be [N,1,H,W]")
        src = np.ascontiguousarray(np.asarray(src_hw, np.float64).reshape(n, 2))
        boxes = torch.empty((n, max_candidates, 8), dtype=torch.float32, device=prob.device)
        counts = torch.empty((n,), dtype=torch.int32, device=prob.device)
        ovf = C.c_int32(0)
        fn = self._lib.dv_db_boxes if variant == "db_pp" else self._lib.dv_db_boxes_dbnet
        check(fn(self._h, _ptr(prob), n, h, w, src.ctypes.data_as(C.POINTER(C.c_double)), float(thresh),
                 float(box_thresh), float(unclip_ratio), int(max_candidates), _ptr(boxes), _ptr(counts),
                 C.byref(ovf) if check_overflow else None), self._h, "dv_db_boxes")
        if check_overflow:
            return boxes, counts, int(ovf.value)
        return boxes, counts

    LORE_MEAN = (0.408, 0.447, 0.470)  # TableLorePreProcessor.process (lore/processer_lore.py:67-70)
    LORE_STD = (0.289, 0.274, 0.278)

    def lore_detect_forward(self, x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """fp32 NCHW [N,3,H,W] (cuda, pre-processed) -> packed head maps fp32 [N,H/4,W/4,24] (hm after sigmoid)."""
        x = _require_cuda(x, torch.float32, "x")
        n, c, h, w = x.shape
        if c != 3:
            raise ValueError("lore expects 3 channels")
        if out is None:
            out = torch.empty((n, h // 4, w // 4, 24), dtype=torch.float32, device=x.device)
        check(self._lib.dv_lore_detect_forward(self._h, _ptr(x), n, h, w, _ptr(out)), self._h, "dv_lore_detect_forward")
        return out

    def lore_detect_forward_u8(self, img: torch.Tensor, out: Optional[torch.Tensor] = None, flip: bool = False) -> torch.Tensor:
        """uint8 HWC [N,H,W,3] (cuda; the warpAffine output) -> packed head maps; normalisation fused on the device."""
        img = _require_cuda(img, torch.uint8, "img")
        n, h, w, c = img.shape
        if c != 3:
            raise ValueError("lore expects HWC images with 3 channels")
        if out is None:
            out = torch.empty((n, h // 4, w // 4, 24), dtype=torch.float32, device=img.device)
        mean = (C.c_float * 3)(*self.LORE_MEAN)
        std = (C.c_float * 3)(*self.LORE_STD)
        check(self._lib.dv_lore_detect_forward_u8(self._h, _ptr(img), n, h, w, mean, std, int(flip), _ptr(out)), self._h,
              "dv_lore_detect_forward_u8")
        return out

    def lore_cell_features(self, dec: dict, max_rows: int, check_overflow: bool = False):
        """Output of lore_decode -> (logi_feat fp32 [max_rows,256], offsets int32 [N+1]) from the resident feature map."""
        n, k = dec["ax_idx"].shape
        dev = dec["ax_idx"].device
        feat = torch.zeros((max_rows, 256), dtype=torch.float32, device=dev)
        offsets = torch.zeros((n + 1,), dtype=torch.int32, device=dev)
        ovf = C.c_int32(0)
        check(self._lib.dv_lore_cell_features(self._h, n, k, int(max_rows), _ptr(dec["counts"]), _ptr(dec["ax_idx"]), _ptr(dec["cr_idx"]),
                                              _ptr(feat), _ptr(offsets), C.byref(ovf) if check_overflow else None), self._h,
              "dv_lore_cell_features")
        if check_overflow and ovf.value:
            raise DocVisionError(f"lore_cell_features: {ovf.value} cells exceed max_rows={max_rows}")
        return feat, offsets

    def lore_add_position_embeddings(self, feat: torch.Tensor, dec: dict, offsets: torch.Tensor) -> torch.Tensor:
        """wiz_2dpe configurations (ptn / wireless): adds x / y position embeddings of the decode's integer position features to
        the packed cell features, in place (on the "lore_processor" handle)."""
        feat = _require_cuda(feat, torch.float32, "feat")
        offsets = _require_cuda(offsets, torch.int32, "offsets")
        dets = _require_cuda(dec["dets_feat"], torch.int32, "dets_feat")
        counts = _require_cuda(dec["counts"], torch.int32, "counts")
        n, k = int(dets.shape[0]), int(dets.shape[1])
        check(self._lib.dv_lore_add_position_embeddings(self._h, _ptr(feat), int(feat.shape[0]), _ptr(dets), _ptr(counts), _ptr(offsets), n, k),
              self._h, "dv_lore_add_position_embeddings")
        return feat

    def lore_process_forward(self, feat: torch.Tensor, offsets: torch.Tensor):
        """feat fp32 [max_rows,256], offsets int32 [N+1] (device) -> (logic [max_rows,4], stacked [max_rows,4]) fp32."""
        feat = _require_cuda(feat, torch.float32, "feat")
        offsets = _require_cuda(offsets, torch.int32, "offsets")
        rows = feat.shape[0]
        n = offsets.shape[0] - 1
        logic = torch.zeros((rows, 4), dtype=torch.float32, device=feat.device)
        stacked = torch.zeros((rows, 4), dtype=torch.float32, device=feat.device)
        check(self._lib.dv_lore_process_forward(self._h, _ptr(feat), rows, _ptr(offsets[n:]), _ptr(offsets), n, _ptr(logic),
                                                _ptr(stacked)), self._h, "dv_lore_process_forward")
        return logic, stacked

    def centernet_decode(self, hm: torch.Tensor, reg, c2v, v2c, inv_affine, K: int = 1000, MK: int = 4000, score_threshold: float = 0.3):
        """CenterNet head maps (four NCHW tensors, or one packed NHWC [N,H,W,24] tensor as `hm` with the others None) ->
        (polygons fp32 [N,K,8] in source pixels, counts int32 [N])."""
        hm = _require_cuda(hm, torch.float32, "hm")
        if reg is None:
            n, h, w, c = hm.shape
            layout = 1
        else:
            reg, c2v, v2c = (_require_cuda(t, torch.float32, nm) for t, nm in ((reg, "reg"), (c2v, "c2v"), (v2c, "v2c")))
            n, c, h, w = hm.shape
            layout = 0
        tr = np.ascontiguousarray(np.asarray(inv_affine, np.float64).reshape(n, 6))
        polygons = torch.empty((n, K, 8), dtype=torch.float32, device=hm.device)
        counts = torch.empty((n,), dtype=torch.int32, device=hm.device)
        check(self._lib.dv_centernet_decode(self._h, _ptr(hm), _ptr(reg), _ptr(c2v), _ptr(v2c), layout, n, h, w,
                                            tr.ctypes.data_as(C.POINTER(C.c_double)), int(K), int(MK), float(score_threshold), _ptr(polygons),
                                            _ptr(counts), None), self._h, "dv_centernet_decode")
        return polygons, counts

    def lore_decode(self, hm: torch.Tensor, reg: Optional[torch.Tensor], wh: Optional[torch.Tensor], st: Optional[torch.Tensor],
                    inv_affine, K: int = 3000, MK: int = 5000, wiz_rev: bool = True, vis_thresh: float = 0.2,
                    check_overflow: bool = False):
        """Lore head maps -> sorted cells.  Either four NCHW fp32 tensors (hm AFTER sigmoid [N,2,H,W], reg [N,2,H,W],
        wh [N,8,H,W], st [N,8,H,W]) or one packed NHWC [N,H,W,24] tensor as `hm` with reg = wh = st = None.
        inv_affine: [N,2,3] float64 (host).  Returns a dict of device tensors: polygons [N,K,8], scores [N,K],
        dets_feat [N,K,8] int32, ax_idx [N,K] int32, cr_idx [N,K,4] int32, counts [N] int32, rows [N] int32."""
        hm = _require_cuda(hm, torch.float32, "hm")
        if reg is None:
            n, h, w, c = hm.shape
            if c != 24:
                raise ValueError("packed Lore maps must be [N,H,W,24]")
            layout = 1
        else:
            reg, wh, st = (_require_cuda(t, torch.float32, nm) for t, nm in ((reg, "reg"), (wh, "wh"), (st, "st")))
            n, c, h, w = hm.shape
            if c != 2 or tuple(reg.shape) != (n, 2, h, w) or tuple(wh.shape) != (n, 8, h, w) or tuple(st.shape) != (n, 8, h, w):
                raise ValueError("Lore maps must be hm [N,2,H,W], reg [N,2,H,W], wh [N,8,H,W], st [N,8,H,W]")
            layout = 0
        tr = np.ascontiguousarray(np.asarray(inv_affine, np.float64).reshape(n, 6))
        dev = hm.device
        out = {
            "polygons": torch.empty((n, K, 8), dtype=torch.float32, device=dev),
            "scores": torch.empty((n, K), dtype=torch.float32, device=dev),
            "dets_feat": torch.empty((n, K, 8), dtype=torch.int32, device=dev),
            "ax_idx": torch.empty((n, K), dtype=torch.int32, device=dev),
            "cr_idx": torch.empty((n, K, 4), dtype=torch.int32, device=dev),
            "counts": torch.empty((n,), dtype=torch.int32, device=dev),
            "rows": torch.empty((n,), dtype=torch.int32, device=dev),
        }
        ovf = C.c_int32(0)
        check(self._lib.dv_lore_decode(self._h, _ptr(hm), _ptr(reg), _ptr(wh), _ptr(st), layout, n, h, w,
                                       tr.ctypes.data_as(C.POINTER(C.c_double)), int(K), int(MK), int(bool(wiz_rev)), float(vis_thresh),
                                       _ptr(out["polygons"]), _ptr(out["scores"]), _ptr(out["dets_feat"]), _ptr(out["ax_idx"]),
                                       _ptr(out["cr_idx"]), _ptr(out["counts"]), _ptr(out["rows"]),
                                       C.byref(ovf) if check_overflow else None), self._h, "dv_lore_decode")
        if check_overflow:
            out["overflow"] = int(ovf.value)
        return out

    def lore_gather_logi(self, ax: torch.Tensor, cr: torch.Tensor, dec: dict) -> torch.Tensor:
        """Dense ax / cr [N,C,H,W] fp32 + the output of lore_decode -> logi_feat [N,K,C] fp32 (rows < counts[n] written)."""
        ax = _require_cuda(ax, torch.float32, "ax")
        cr = _require_cuda(cr, torch.float32, "cr")
        n, c, h, w = ax.shape
        k = dec["ax_idx"].shape[1]
        out = torch.zeros((n, k, c), dtype=torch.float32, device=ax.device)
        check(self._lib.dv_lore_gather_logi(self._h, _ptr(ax), _ptr(cr), n, c, h, w, k, _ptr(dec["counts"]), _ptr(dec["ax_idx"]),
                                            _ptr(dec["cr_idx"]), _ptr(out)), self._h, "dv_lore_gather_logi")
        return out

    PICODET_MEAN = (0.485, 0.456, 0.406)  # PicodetConfig norm_mean / norm_std (picodet/configuration_picodet.py:50-51)
    PICODET_STD = (0.229, 0.224, 0.225)
    PICODET_STRIDES = (8, 16, 32, 64)

    def _picodet_outputs(self, n, h, w, dev):
        c = int(self._lib.dv_picodet_num_classes(self._h))
        hw = [-(-h // s) * -(-w // s) for s in self.PICODET_STRIDES]
        scores = [torch.empty((n, m, c), dtype=torch.float32, device=dev) for m in hw]
        dfl = [torch.empty((n, m, 32), dtype=torch.float32, device=dev) for m in hw]
        return scores, dfl, (C.c_void_p * 4)(*[t.data_ptr() for t in scores]), (C.c_void_p * 4)(*[t.data_ptr() for t in dfl])

    def picodet_forward(self, x: torch.Tensor):
        """fp32 NCHW [N,3,H,W] (cuda, pre-processed) -> (scores[4] fp32 [N,HW_l,C], dfl[4] fp32 [N,HW_l,32])."""
        x = _require_cuda(x, torch.float32, "x")
        n, c, h, w = x.shape
        if c != 3:
            raise ValueError("picodet expects 3 channels")
        scores, dfl, sp, dp = self._picodet_outputs(n, h, w, x.device)
        check(self._lib.dv_picodet_forward(self._h, _ptr(x), n, h, w, sp, dp), self._h, "dv_picodet_forward")
        return scores, dfl

    def picodet_forward_u8(self, img: torch.Tensor, flip: bool = True):
        """uint8 HWC [N,H,W,3] (cuda; the cv2.resize output) -> (scores[4], dfl[4]); flip + normalisation fused."""
        img = _require_cuda(img, torch.uint8, "img")
        n, h, w, c = img.shape
        if c != 3:
            raise ValueError("picodet expects HWC images with 3 channels")
        scores, dfl, sp, dp = self._picodet_outputs(n, h, w, img.device)
        mean = (C.c_float * 3)(*self.PICODET_MEAN)
        std = (C.c_float * 3)(*self.PICODET_STD)
        check(self._lib.dv_picodet_forward_u8(self._h, _ptr(img), n, h, w, mean, std, float(np.float32(1.0 / 255.0)), int(flip), sp, dp),
              self._h, "dv_picodet_forward_u8")
        return scores, dfl

    # ------------------------------------------------------------------ PP-OCR recogniser (model kind "pp_rec")
    def rec_time_steps(self, height: int, width: int) -> int:
        return int(self._lib.dv_rec_time_steps(self._h, int(height), int(width)))

    @property
    def rec_num_classes(self) -> int:
        return int(self._lib.dv_rec_num_classes(self._h))

    def _rec_outputs(self, n, h, w, dev, return_probs):
        t, c = self.rec_time_steps(h, w), self.rec_num_classes
        if t <= 0:
            raise ValueError(f"rec_forward: a {h}x{w} input is too small for the network")
        ids = torch.empty((n, t), dtype=torch.int32, device=dev)
        maxp = torch.empty((n, t), dtype=torch.float32, device=dev)
        probs = torch.empty((n, t, c), dtype=torch.float32, device=dev) if return_probs else None
        return ids, maxp, probs

    def rec_forward(self, x: torch.Tensor, return_probs: bool = False):
        """fp32 NCHW [N,3,48,W] (cuda; PPOcrRecPreProcessor's batch) -> (ids int32 [N,T], maxp fp32 [N,T][, probs fp32 [N,T,C]])."""
        x = _require_cuda(x, torch.float32, "x")
        n, c, h, w = x.shape
        if c != 3:
            raise ValueError("rec_forward expects [N,3,H,W]")
        ids, maxp, probs = self._rec_outputs(n, h, w, x.device, return_probs)
        check(self._lib.dv_rec_forward(self._h, _ptr(x), n, h, w, _ptr(probs), _ptr(ids), _ptr(maxp)), self._h, "dv_rec_forward")
        return (ids, maxp, probs) if return_probs else (ids, maxp)

    def rec_forward_u8(self, crops: torch.Tensor, widths: Optional[torch.Tensor] = None, return_probs: bool = False):
        """uint8 HWC [N,48,W,3] (cuda; resized crops, left aligned, valid up to widths[n]) -> as rec_forward; the normalisation
        (x / 255 - 0.5) / 0.5 and the zero padding beyond each width are fused into the first kernel."""
        crops = _require_cuda(crops, torch.uint8, "crops")
        n, h, w, c = crops.shape
        if c != 3:
            raise ValueError("rec_forward_u8 expects [N,H,W,3]")
        if widths is not None:
            widths = _require_cuda(widths, torch.int32, "widths")
            if widths.numel() != n:
                raise ValueError("one width per crop")
        ids, maxp, probs = self._rec_outputs(n, h, w, crops.device, return_probs)
        check(self._lib.dv_rec_forward_u8(self._h, _ptr(crops), _ptr(widths), n, h, w, _ptr(probs), _ptr(ids), _ptr(maxp)), self._h,
              "dv_rec_forward_u8")
        return (ids, maxp, probs) if return_probs else (ids, maxp)

    def cls_forward(self, x: torch.Tensor, return_probs: bool = False):
        """PULC classifier (model kind "pplcnet_cls"): fp32 NCHW [N,3,H,W] (cuda) -> logits fp32 [N,C] (+ softmax probabilities)."""
        x = _require_cuda(x, torch.float32, "x")
        n, c, h, w = x.shape
        if c != 3:
            raise ValueError("cls_forward expects [N,3,H,W]")
        nc = self.rec_num_classes
        logits = torch.empty((n, nc), dtype=torch.float32, device=x.device)
        probs = torch.empty((n, nc), dtype=torch.float32, device=x.device) if return_probs else None
        check(self._lib.dv_cls_forward(self._h, _ptr(x), n, h, w, _ptr(logits), _ptr(probs)), self._h, "dv_cls_forward")
        return (logits, probs) if return_probs else logits

    def picodet_decode(self, scores, dfl, org_hw, scale_factor, in_hw=(800, 608), strides=(8, 16, 32, 64), score_threshold: float = 0.5,
                       nms_threshold: float = 0.5, nms_top_k: int = 1000, keep_top_k: int = 100):
        """scores[l] fp32 [N,HW_l,C], dfl[l] fp32 [N,HW_l,4*(reg_max+1)] (cuda, 4 levels); org_hw [N,2] (h, w), scale_factor [N,2]
        (ratio_h, ratio_w) -> (boxes float64 [N, C*keep_top_k, 6] rows (class, score, x1, y1, x2, y2), counts int32 [N])."""
        if len(scores) != 4 or len(dfl) != 4:
            raise ValueError("picodet_decode expects 4 levels")
        scores = [_require_cuda(t, torch.float32, "scores") for t in scores]
        dfl = [_require_cuda(t, torch.float32, "dfl") for t in dfl]
        n, _, c = scores[0].shape
        reg_max = dfl[0].shape[-1] // 4 - 1
        for lvl, st in enumerate(strides):
            hw = -(-in_hw[0] // st) * -(-in_hw[1] // st)
            if tuple(scores[lvl].shape) != (n, hw, c) or tuple(dfl[lvl].shape) != (n, hw, 4 * (reg_max + 1)):
                raise ValueError(f"level {lvl}: expected [{n},{hw},*] tensors")
        cap = c * keep_top_k
        out = torch.zeros((n, cap, 6), dtype=torch.float64, device=scores[0].device)
        counts = torch.zeros((n,), dtype=torch.int32, device=scores[0].device)
        sp = (C.c_void_p * 4)(*[t.data_ptr() for t in scores])
        dp = (C.c_void_p * 4)(*[t.data_ptr() for t in dfl])
        org = np.ascontiguousarray(np.asarray(org_hw, np.float32).reshape(n, 2))
        sf = np.ascontiguousarray(np.asarray(scale_factor, np.float32).reshape(n, 2))
        st = (C.c_int * 4)(*strides)
        check(self._lib.dv_picodet_decode(self._h, sp, dp, n, c, reg_max, st, int(in_hw[0]), int(in_hw[1]),
                                          org.ctypes.data_as(C.POINTER(C.c_float)), sf.ctypes.data_as(C.POINTER(C.c_float)),
                                          float(score_threshold), float(nms_threshold), int(nms_top_k), int(keep_top_k), cap, _ptr(out),
                                          _ptr(counts)), self._h, "dv_picodet_decode")
        return out, counts

    def ctc_greedy(self, probs: torch.Tensor, blank: int = 0, return_raw: bool = False):
        """[B,T,C] fp32 (cuda) -> (ids [B,T] int32 left-packed / -1 padded, len [B] int32, conf [B] fp32)."""
        probs = _require_cuda(probs, torch.float32, "probs")
        b, t, c = probs.shape
        dev = probs.device
        ids = torch.empty((b, t), dtype=torch.int32, device=dev)
        ln = torch.empty((b,), dtype=torch.int32, device=dev)
        conf = torch.empty((b,), dtype=torch.float32, device=dev)
        raw_ids = torch.empty((b, t), dtype=torch.int32, device=dev) if return_raw else None
        raw_max = torch.empty((b, t), dtype=torch.float32, device=dev) if return_raw else None
        check(self._lib.dv_ctc_greedy(self._h, _ptr(probs), b, t, c, blank, _ptr(ids), _ptr(ln), _ptr(conf),
                                      _ptr(raw_ids), _ptr(raw_max)), self._h, "dv_ctc_greedy")
        if return_raw:
            return ids, ln, conf, raw_ids, raw_max
        return ids, ln, conf

    # ------------------------------------------------------------------ operator level (kernel parity tests)
    def nchw_to_nhwc_f16(self, x: torch.Tensor) -> torch.Tensor:
        x = _require_cuda(x, torch.float32, "x")
        n, c, h, w = x.shape
        out = torch.empty((n, h, w, c), dtype=torch.float16, device=x.device)
        check(self._lib.dv_nchw_f32_to_nhwc_f16(self._h, _ptr(x), n, c, h, w, _ptr(out)), self._h)
        return out

    def nhwc_f16_to_nchw(self, x: torch.Tensor) -> torch.Tensor:
        x = _require_cuda(x, torch.float16, "x")
        n, h, w, c = x.shape
        out = torch.empty((n, c, h, w), dtype=torch.float32, device=x.device)
        check(self._lib.dv_nhwc_f16_to_nchw_f32(self._h, _ptr(x), n, c, h, w, _ptr(out)), self._h)
        return out

    def conv2d_nhwc(self, x_nhwc: torch.Tensor, w_packed: torch.Tensor, bias: Optional[torch.Tensor], cout: int,
                    ksize: int, stride: int, pad: int, residual: Optional[torch.Tensor] = None, act: int = 0):
        x_nhwc = _require_cuda(x_nhwc, torch.float16, "x")
        w_packed = _require_cuda(w_packed, torch.float16, "w")
        n, h, w, cin = x_nhwc.shape
        cin_pad = w_packed.shape[1] // (ksize * ksize)
        ho = (h + 2 * pad - ksize) // stride + 1
        wo = (w + 2 * pad - ksize) // stride + 1
        out = torch.empty((n, ho, wo, cout), dtype=torch.float16, device=x_nhwc.device)
        if bias is not None:
            bias = _require_cuda(bias, torch.float32, "bias")
        if residual is not None:
            residual = _require_cuda(residual, torch.float16, "residual")
        check(self._lib.dv_conv2d_nhwc_f16(self._h, _ptr(x_nhwc), n, h, w, cin, _ptr(w_packed), cin_pad, _ptr(bias),
                                           cout, ksize, stride, pad, _ptr(residual), act, _ptr(out)), self._h,
              "dv_conv2d_nhwc_f16")
        return out
