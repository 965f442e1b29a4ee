"""The model-calling methods of the reference's orchestrator, OcrSystemTask (model/ocr_pdf/ocr_system_task.py), over the
b200 predictors -- the callers either side of the hot path: text_detection (:146-166), text_recognition
(:300-330), layout_analysis (:214-225) and the table loop of table_structure_detection (:179-198).  Same names, arguments and
return conventions ``(result, metric)``; what changes is where the pixels live: a page goes to the device once, every
text-line crop and every table crop is cut from the resident page by a kernel (OcrRecognitionTask.recognize_page,
OcrTableStructureTask.recognize_tables) and only boxes, token ids and cells come back.

Everything else of the orchestrator (PDF parsing, cell / text matching, HTML export, file outputs) stays the reference's:
these methods return exactly what its next steps consume.
"""
from __future__ import annotations

import contextlib
import time
from typing import Any, Dict, List, Optional, Tuple

import numpy as np
import torch

from .predictors import _h2d, order_point, order_points_batch, sort_det_boxes

__all__ = ["OcrSystemTask", "get_layout_by_type", "match_table_cell_and_text_cell"]


def get_layout_by_type(layout_result: List[Dict[str, Any]], label: str = "figure", score_threshold: float = 0.8) -> List[Dict[str, Any]]:
    """TableProcessUtils.get_layout_by_type (model/pdf_table/table_common.py:1286-1299): the layout boxes of one label at or
    above the score threshold, sorted by their top edge (stable)."""
    results = [item for item in layout_result if item["label"].lower() == label.lower() and item["score"] >= score_threshold]
    results.sort(key=lambda x: x["bbox"][1])
    return results


def match_table_cell_and_text_cell(engine, table_cells, text_bboxs) -> Dict[int, List[int]]:
    """The matching half of OcrTableToHtmlTask.match_table_cell_and_text_cell (ocr_pdf/ocr_table_to_html_task.py:178-207): every
    text box goes to the table cell find_top1_mach_box picks (:48-77) -- on the device (dv_match_cells: one warp per text box,
    float64 arithmetic in the reference's operation order, identical indices).  table_cells / text_bboxs: sequences of
    [x1, y1, x2, y2] (Cell.to_bbox / OcrCell.to_bbox).  Returns the reference's ``matched`` dict {cell index: [text indices in
    input order]} (insertion-ordered by first match, like the reference's loop); text merging and the HTML export stay the
    reference's."""
    matched: Dict[int, List[int]] = {}
    if len(text_bboxs) == 0:
        return matched
    dev = torch.device("cuda", engine.device)
    t = torch.as_tensor(np.asarray(text_bboxs, dtype=np.float64).reshape(-1, 4)).to(dev)
    c = torch.as_tensor(np.asarray(table_cells, dtype=np.float64).reshape(-1, 4)).to(dev)
    top1 = engine.match_cells(t, c).cpu().numpy()
    for index, cell in enumerate(top1.tolist()):
        matched.setdefault(int(cell), []).append(index)
    return matched


class OcrSystemTask:
    """Holds already constructed predictors (any may be None; using a missing one raises RuntimeError, where the reference
    would lazily construct its default model)."""

    def __init__(self, text_detector=None, text_recognizer=None, table_structure_recognizer=None, layout_detector=None,
                 table_stream: bool = True):
        # table_stream: predict_pages runs the table-structure branch on its own CUDA stream beside detect -> recognise (the
        # branches share only the resident pages); False keeps everything on the caller's stream
        self.table_stream = table_stream
        self._side: Optional[torch.cuda.Stream] = None
        self._chain: Optional[torch.cuda.Stream] = None
        self._copy: Optional[torch.cuda.Stream] = None
        self.text_detector = text_detector
        self.text_recognizer = text_recognizer
        self.table_structure_recognizer = table_structure_recognizer
        self.layout_detector = layout_detector

    @staticmethod
    def _need(task, name: str):
        if task is None:
            raise RuntimeError(f"OcrSystemTask: no {name} was given")
        return task

    def page_to_device(self, image_full) -> torch.Tensor:
        """uint8 HWC ndarray (or cuda tensor) -> the resident page all crops are cut from."""
        if isinstance(image_full, torch.Tensor):
            return image_full
        task = self.text_recognizer or self.table_structure_recognizer or self.text_detector or self.layout_detector
        dev = torch.device("cuda", self._need(task, "predictor").device)
        return torch.from_numpy(np.ascontiguousarray(image_full)).to(dev)

    # ------------------------------------------------------------------ the page loop, batched
    @staticmethod
    def _upload(pages, dev: torch.device) -> torch.Tensor:
        if isinstance(pages, np.ndarray) and pages.ndim == 4:
            return _h2d(torch.from_numpy(pages), dev)
        return torch.stack([_h2d(torch.from_numpy(np.ascontiguousarray(p)), dev) for p in pages])

    def predict_stream(self, batches, **kwargs):
        """The throughput form of ``predict_pages``: a generator over an iterable of page batches that yields what
        ``predict_pages(batch, **kwargs)`` returns for each, with batch i + 1 uploaded by a copy stream while batch i computes
        (give it views of pinned memory: the copy engine then works beside the SMs and the step no longer starts with an idle
        device waiting for its 2.8 MB per page) and the results of batch i - 1 collected (read-back, string building, result dicts)
        after batch i has been enqueued, so the host's share of a batch overlaps the device's share of the next.  Two batches are
        in flight at a time; the results are those of ``predict_pages``, one batch late."""
        det = self._need(self.text_detector, "text_detector")
        dev = torch.device("cuda", det.device)
        if self._copy is None or self._copy.device != dev:
            self._copy = torch.cuda.Stream(device=dev)

        def start(pages):
            with torch.cuda.stream(self._copy):
                batch = self._upload(pages, dev)
                ev = torch.cuda.Event()
                ev.record()
            return batch, ev

        keep = bool(kwargs.pop("keep_device_record", False))
        it = iter(batches)
        nxt = next(it, None)
        up = start(nxt) if nxt is not None else None
        pending = None
        while up is not None:
            cur, nxt = up, next(it, None)
            up = start(nxt) if nxt is not None else None
            # everything of batch i is enqueued (its recognition pass follows the host's look at its detected boxes) BEFORE the
            # results of batch i - 1 are collected: the host's result building and the read-back of batch i - 1 then run while
            # the device works on batch i.  Every run owns its result tensors and the engines' internal buffers are reused in
            # stream order, so the two batches in flight do not meet.
            state = self.predict_pages(None, _uploaded=cur, _defer=True, **kwargs)
            if pending is not None:
                yield self._finish_batch(pending, keep)
            pending = state
        if pending is not None:
            yield self._finish_batch(pending, keep)

    def predict_pages(self, pages, layout_tables=None, det_kwargs: Optional[Dict[str, Any]] = None,
                      keep_device_record: bool = False, _uploaded=None, _defer: bool = False) -> List[Dict[str, Any]]:
        """The reference's per-page sequence (cli/main.py:116-144 -> ocr_system_task.py:549-734: layout_analysis,
        text_detection, text_recognition, table_structure_detection) for a BATCH of equally sized pages, with the host steps
        of one stage overlapped with the device work of another:

            upload pages once -> [GPU] layout + detection -> host: layout records, table boxes -> [GPU] table structure (all
            tables of all pages, cut from the resident pages) || host: reading-order sort + order_point of the detected boxes
            -> [GPU] crop + recognise (all quads of all pages) || host: nothing left -> collect texts, cells.

        pages: uint8 [P,H,W,3] ndarray (may be a view of pinned memory: the upload is then asynchronous) or a list of equally
        sized HWC ndarrays.  layout_tables: optional per-page lists of table boxes [x1,y1,x2,y2] that replace the layout
        detector's own "table" rows (what the reference passes in as ``layout_result``).  det_kwargs: passed to the text
        detector's ``_run_model`` (e.g. ``prob_override``).  Returns one dict per page:
        {"layout": [...], "det": float64 [n,8] in reading order, "ocr": [{"index","text","bbox"}], "tables": [[bbox, result]]}.
        With keep_device_record the packed device-side results of the batch stay in ``self.device_record`` (the record the
        multi-GPU all-gather exchanges, sharding.FIELDS)."""
        det, rec = self._need(self.text_detector, "text_detector"), self._need(self.text_recognizer, "text_recognizer")
        dev = torch.device("cuda", det.device)
        if _uploaded is not None:  # predict_stream: the batch was put on its way by the copy stream while the previous one computed
            batch, uploaded = _uploaded
            torch.cuda.current_stream(dev).wait_event(uploaded)
            batch.record_stream(torch.cuda.current_stream(dev))
        else:
            batch = self._upload(pages, dev)
            uploaded = torch.cuda.Event()
            uploaded.record()  # all the table branch needs from the main stream
        n_pages = int(batch.shape[0])
        page_list = list(batch)
        tsr = self.table_structure_recognizer
        tsr_run = None
        tables_flat: List[Dict[str, Any]] = []

        def launch_tables(layouts):
            """All tables of the batch, cut from the resident pages, on the side stream when there is one (latency-bound kernels
            of one branch then share the SMs with the other branch's: 55.8 -> 53.4 ms per 32-page step on the B200)."""
            for p in range(n_pages):
                if layout_tables is not None:
                    boxes = [{"bbox": b} for b in layout_tables[p]]
                else:
                    boxes = get_layout_by_type(layout_result=layouts[p], label="table", score_threshold=0.2)
                tables_flat.extend({"bbox": t["bbox"], "page": p} for t in boxes)
            if not self.table_stream:
                return tsr.launch_tables(batch, tables_flat)
            if self._side is None or self._side.device != dev:
                self._side = torch.cuda.Stream(device=dev)
            with torch.cuda.stream(self._side):
                self._side.wait_event(uploaded)
                run = tsr.launch_tables(batch, tables_flat)
            batch.record_stream(self._side)
            return run

        # The layout -> detection -> recognition chain runs on a HIGH-priority stream when the tables have their own: the chain
        # then finishes first and the host's share of the text results (dictionary lookups, string building: ~5 ms per 32 pages)
        # overlaps the rest of the table branch instead of following the whole device step.
        chain = contextlib.nullcontext()
        if self.table_stream and tsr is not None:
            if self._chain is None or self._chain.device != dev:
                self._chain = torch.cuda.Stream(device=dev, priority=-1)
            self._chain.wait_event(uploaded)
            batch.record_stream(self._chain)
            chain = torch.cuda.stream(self._chain)
        with chain:
            # ---- stage 1 (GPU): layout + detection enqueued back to back
            lay_run = None
            if self.layout_detector is not None:
                lay_run = self.layout_detector._run_model(self.layout_detector._preprocess(page_list))
            det_run = det._run_model(det._preprocess(page_list), **(det_kwargs or {}))
            # with the table boxes given, the table branch does not wait for the layout results: its host preparation (crop
            # rects, affine matrices) runs while the device already works on stage 1
            if tsr is not None and layout_tables is not None:
                tsr_run = launch_tables(None)
            # ---- host: layout records -> table boxes; stage 2 (GPU): table structure
            layouts = self.layout_detector._postprocess(lay_run) if lay_run is not None else [[] for _ in range(n_pages)]
            if tsr is not None and layout_tables is None:
                tsr_run = launch_tables(layouts)
            # ---- host (while the tables run): reading order + corner order of the detected boxes; stage 3 (GPU): recognition
            dets = [sort_det_boxes(d) if len(d) else np.zeros((0, 8)) for d in det._postprocess(det_run)]
            pts = [order_points_batch(d) for d in dets]
            rec_run = rec.launch_pages(batch, pts)
        state = {"batch": batch, "n_pages": n_pages, "layouts": layouts, "dets": dets, "pts": pts, "det_run": det_run, "rec_run": rec_run,
                 "tsr_run": tsr_run, "tables_flat": tables_flat}
        return state if _defer else self._finish_batch(state, keep_device_record)

    def _finish_batch(self, st: Dict[str, Any], keep_device_record: bool) -> List[Dict[str, Any]]:
        """Second half of ``predict_pages``: waits for the copies and builds the per-page results -- the texts first (their chain
        ends first), the tables last."""
        n_pages, pts = st["n_pages"], st["pts"]
        texts = self.text_recognizer.collect_pages(st["rec_run"])
        tables = self.table_structure_recognizer.collect_tables(st["tsr_run"]) if st["tsr_run"] is not None else []
        out = []
        for p in range(n_pages):
            ocr = [{"index": i + 1, "text": "" if t is None else t, "bbox": q} for i, (t, q) in enumerate(zip(texts[p], pts[p]))]
            out.append({"layout": st["layouts"][p], "det": st["dets"][p], "ocr": ocr, "tables": []})
        for tb, res in zip(st["tables_flat"], tables):
            out[tb["page"]]["tables"].append(res)
        if keep_device_record:
            self.device_record = self._device_record(st["det_run"], st["rec_run"], st["tsr_run"], n_pages)
        return out

    @staticmethod
    def _device_record(det_run, rec_run, tsr_run, n_pages: int) -> Dict[str, torch.Tensor]:
        """The packed per-batch results that stay on the device (sharding.FIELDS / TABLE_FIELDS): boxes + counts of the pages
        (single shape group), collapsed token ids + lengths of all crops, cells + counts + logical coordinates of all tables."""
        _, boxes, counts, _, _ = det_run["pending"][0]
        rec = {"boxes": boxes[:, :64].contiguous(), "box_counts": counts}
        if rec_run.get("event") is not None:
            rec["ids"], rec["id_lens"] = rec_run["ids_dev"], rec_run["len_dev"]
        else:
            rec["ids"] = torch.zeros((0, 201), dtype=torch.int32, device=boxes.device)
            rec["id_lens"] = torch.zeros((0,), dtype=torch.int32, device=boxes.device)
        if tsr_run is not None and "logi" in tsr_run["dev"]:
            d = tsr_run["dev"]
            # the logical coordinates are packed by the per-image offsets: re-pad them to 256 rows per table (index plumbing)
            idx = d["offsets"][:-1].long()[:, None] + torch.arange(256, device=d["logi"].device)[None, :]
            rec["cells"] = d["polygons"][:, :256].contiguous()
            rec["cell_counts"] = d["counts"]
            rec["cell_logi"] = d["logi"][idx.clamp_(max=int(d["logi"].shape[0]) - 1)]
        return rec

    def text_detection(self, image) -> Tuple[np.ndarray, Dict[str, float]]:
        """:146-166: detect, then sort the boxes into reading order.  Returns (float64 [n,8], metric)."""
        start = time.time()
        det_result = self._need(self.text_detector, "text_detector")(image)[0]
        use_time = time.time() - start
        return sort_det_boxes(det_result), {"use_time": use_time}

    def text_recognition(self, det_result, image_full) -> Tuple[List[Dict[str, Any]], Dict[str, Any]]:
        """:300-330: per box order_point -> crop_image -> recognise; here all crops of the page in one device pass.  A crop the
        reference fails on (empty crop: it logs the exception and keeps "") yields "" as there.  Returns (list of
        {"index", "text", "bbox"}, metric)."""
        rec = self._need(self.text_recognizer, "text_recognizer")
        start = time.time()
        det_result = np.asarray(det_result)
        pts = [order_point(det_result[i]) for i in range(det_result.shape[0])]
        texts = rec.recognize_page(self.page_to_device(image_full), pts) if pts else []
        output = [{"index": i + 1, "text": "" if t is None else t, "bbox": p} for i, (t, p) in enumerate(zip(texts, pts))]
        total_use_time = time.time() - start
        if not output:  # the reference divides by len(use_times) here (:322) -> ZeroDivisionError on a page without boxes
            raise ZeroDivisionError("text_recognition: no detected boxes")
        metric = {"use_time": total_use_time, "avg_use_time": total_use_time / len(output), "total": len(output)}
        return output, metric

    def layout_analysis(self, image) -> Tuple[List[Dict[str, Any]], Dict[str, float]]:
        """:214-225."""
        start = time.time()
        layout_result = self._need(self.layout_detector, "layout_detector")(image)[0]
        return layout_result, {"use_time": time.time() - start}

    def table_structure_detection(self, image, image_full=None, layout_result: Optional[List[Dict[str, Any]]] = None):
        """:179-203: with a layout result, every "table" box (score >= 0.2, top to bottom) is cropped from image_full and
        recognised; returns (outputs = [[bbox, one_result], ...] -- the argument of the reference's
        TableProcessUtils.convert_table_sep_to_merge (:199) --, metric).  Without one the recogniser runs on the whole image
        and its first result is returned, as there (:202-203)."""
        tsr = self._need(self.table_structure_recognizer, "table_structure_recognizer")
        start = time.time()
        if layout_result:
            layout_tables = get_layout_by_type(layout_result=layout_result, label="table", score_threshold=0.2)
            page = self.page_to_device(image if image_full is None else image_full)
            outputs = tsr.recognize_tables(page, [{"bbox": t["bbox"]} for t in layout_tables])
            return outputs, {"use_time": time.time() - start}
        result = tsr(image)
        return result[0], {"use_time": time.time() - start}
