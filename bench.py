#!/usr/bin/env python
"""bench.py -- the driver's measurement contract for the pdf_table hot path on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--cascade full|ocr] [--det ppocrv4|dbnet_r18]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

ONE JSON line on rank 0.  The headline (`metric`/`value`/`e2e`) is BASELINE.json's second quantity, pages/s through the
FULL cascade (configs[4] per GPU: PicoDet layout -> DB detect -> text-line recognise -> Lore table structure, 32 synthetic
960x960 pages per GPU, weak scaling); the same line carries, under `blocks`, the metric's first quantity -- text-line
crops/s of the recogniser on configs[3] (4096 crops 32x320 dealt over the N GPUs, strong scaling) -- and Lore on
configs[2] (16 x 1024x1024 table crops per GPU), each with its own device value, e2e, roofline and CPU sample.

* `value`  : K timed steps of the cascade with the uint8 pages already resident in HBM, L2 flushed between steps, CUDA
             events, max over ranks.  With N > 1 the ONE collective of the path -- the all-gather of the packed decoded
             results (boxes, token ids, table cells) -- runs INSIDE every timed step.
* `e2e`    : the same batch through the public API a user calls (`pdf_table_b200.system.OcrSystemTask.predict_pages`
             over the predictor classes): numpy pages in pinned host memory in, Python results (boxes, strings, cells)
             out; upload, every host step of the predictors / orchestrator glue, read-back and (N > 1) the all-gather
             are inside the timed region.
* `roofline`: the dominant kernel's algorithmic FLOPs (or bytes) / its CUDA-event time over the same steps.
* `cpu_baseline` / `--impl reference`: the oracle/ restatement of the same stages (the reference's algorithm in plain
             PyTorch fp32 / numpy) on the host cores, on a bounded sample; rank 0 only, in-line only at N = 1.
"""
from __future__ import annotations

import argparse
import glob
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

# NCCL's init lines ("... rank r nranks N ...") are the driver's evidence that the ranks really form one communicator: log
# them (INFO, INIT subsystem) instead of silencing them, but never on stdout -- rank 0's stdout is exactly one JSON line.
os.environ.setdefault("NCCL_DEBUG", "INFO")
if os.environ["NCCL_DEBUG"].upper() == "INFO":
    os.environ.setdefault("NCCL_DEBUG_SUBSYS", "INIT")
os.environ.setdefault("NCCL_DEBUG_FILE", os.path.join(os.environ.get("TMPDIR", "/tmp"), "nccl.bench.%h.%p.log"))

import numpy as np  # noqa: E402
import torch  # noqa: E402

PAGES_PER_GPU = 32
PAGE_H = PAGE_W = 960
CROPS_PER_PAGE = 40          # synthetic pages plant 30-60 text lines
CTC_T, CTC_C = 40, 97        # PP-OCRv4 en rec head: 48x320 crop -> T=40, C=97
MEAN = (0.485, 0.456, 0.406)
STD = (0.229, 0.224, 0.225)
FALLBACK_PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}
TABLE_BBOX = [24.3, 130.6, 936.4, 830.5]  # the layout box every page's table is cut at
LORE_HM_BIAS = (-0.3, -3.5)  # seeded random weights: shift the Lore heat maps so that ~100 cells / corners per table pass the gates
TABLES_PER_PAGE = 1
SWEEP_CROPS, SWEEP_H, SWEEP_W = 4096, 32, 320   # BASELINE configs[3]
LORE_BATCH = 16                                 # BASELINE configs[2]
PP_REC_CLASSES = 97                             # en dictionary of the PP-OCRv4 recogniser
FULL = True
TWO_STREAMS = os.environ.get("DV_BENCH_STREAMS", "2") == "2"  # the table-structure branch on its own stream, as OcrSystemTask.predict_pages runs it (1: one stream)
DET = "ppocrv4"  # --det: ppocrv4 = PPLCNetV3-0.75 + RSE-FPN + DBHead (the PP-OCRv4 det graph); dbnet_r18 = the in-tree DBModel


def measured_traffic(kernel: str, workload: str):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of `kernel`, from the newest committed ncu pass over THIS
    workload (profiles/*_dram_traffic.json with a matching "workload" key); None when no capture covers it."""
    for f in reversed(sorted(glob.glob(os.path.join(ROOT, "profiles", "*_dram_traffic.json")))):
        try:
            with open(f) as fh:
                d = json.load(fh)
            if d.get("workload", "ocr") != workload:
                continue
            k = d["kernels"].get(kernel)
        except (OSError, ValueError, KeyError):
            continue
        if k:
            return k["dram_read_bytes_per_launch"] + k["dram_write_bytes_per_launch"], "profiles/" + os.path.basename(f)
    return None, None


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            out = {k: float(d[k]) for k in ("hbm_gbs", "bf16_tflops", "bf16_tflops_sustained") if k in d}
            if "hbm_gbs" in out and "bf16_tflops" in out:
                out.setdefault("bf16_tflops_sustained", out["bf16_tflops"])
                return out, "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return dict(FALLBACK_PEAKS), "fallback (B200_PROFILING.md)"


# --------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """Samples nvidia-smi SM clocks / throttle reasons during the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for s in self.samples:
            f = [x.strip() for x in s.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# --------------------------------------------------------------------------------------- synthetic workload
def make_pages(rank: int, n: int) -> np.ndarray:
    from pdf_table_b200 import synth

    # a handful of distinct pages tiled to the batch keeps start-up short; content does not change the work
    distinct = [synth.synthetic_page(rank * 1000 + i, PAGE_H, PAGE_W) for i in range(4)]
    return np.stack([distinct[i % 4] for i in range(n)])


def make_prob_maps(rank: int, n: int) -> np.ndarray:
    """Planted DB probability maps (planted network outputs for post-processor throughput): with seeded
    random weights the detector's own output is texture noise, so the box stage runs on analytic text-line blobs."""
    from pdf_table_b200 import synth

    distinct = [synth.synthetic_prob_map_lines(rank * 1000 + i, PAGE_H, PAGE_W, CROPS_PER_PAGE) for i in range(4)]
    return np.stack([distinct[i % 4] for i in range(n)])[:, None]


def make_crops(rank: int, n: int) -> np.ndarray:
    from pdf_table_b200 import synth

    distinct = [synth.synthetic_text_crop(rank * 100000 + i, 32, 320) for i in range(64)]
    return np.stack([distinct[i % 64] for i in range(n)])


def make_sweep_crops(n: int, first: int = 0, stride: int = 1) -> np.ndarray:
    """Crops first, first + stride, ... of the global configs[3] list (64 distinct synthetic text lines, cycled)."""
    from pdf_table_b200 import synth

    distinct = [synth.synthetic_text_crop(900000 + i, SWEEP_H, SWEEP_W) for i in range(64)]
    return np.stack([distinct[(first + k * stride) % 64] for k in range(n)])


def make_ctc_probs(rank: int, n_crops: int) -> np.ndarray:
    rng = np.random.default_rng(77 + rank)
    logits = rng.standard_normal((n_crops, CTC_T, CTC_C)).astype(np.float32) * 3
    runs = rng.integers(0, CTC_C, size=(n_crops, CTC_T))
    runs[rng.random((n_crops, CTC_T)) < 0.3] = 0
    logits[np.arange(n_crops)[:, None], np.arange(CTC_T)[None, :], runs] += 8
    e = np.exp(logits - logits.max(-1, keepdims=True))
    return (e / e.sum(-1, keepdims=True)).astype(np.float32)


def make_table_crops(rank: int, n: int):
    """1024x1024 table crops already through TableLorePreProcessor's warp (uint8), + inverse affines (configs[2] input)."""
    from pdf_table_b200 import predictors, synth

    pre = [predictors.lore_preprocess(synth.synthetic_page(rank * 1000 + 500 + i, 1024, 1024)) for i in range(4)]
    imgs = np.stack([pre[i % 4][0] for i in range(n)])
    inv = np.stack([predictors.lore_affine([np.float32(pre[i % 4][1][0]), np.float32(pre[i % 4][1][1])], np.float32(pre[i % 4][1][2]), 256, 256, True)
                    for i in range(n)])
    return imgs, inv


def make_layout_pages(pages: np.ndarray) -> np.ndarray:
    import cv2

    return np.stack([cv2.resize(p, (608, 800)) for p in pages])


def synthetic_vocab(n_labels: int):
    """A stand-in vocab file for the recogniser's label mapping (ids 2.. -> characters)."""
    return [chr(0x4E00 + i) for i in range(n_labels - 2)]


# --------------------------------------------------------------------------------------- the B200 arm
class Cascade:
    """Product code only (pdf_table_b200), no oracle imports.  Holds the predictors (the public API of the e2e leg) and
    drives the SAME engine handles at the C-ABI level for the device-resident leg."""

    def __init__(self, rank: int, device: int, full: bool, precision: str = "fp16"):
        from pdf_table_b200 import predictors, synth
        from pdf_table_b200.system import OcrSystemTask

        self.precision = precision  # "fp32x": split-fp16 operand pairs / fp32 buffers in the detector, the recogniser and the Lore detector

        self.full, self.device = full, device
        dev = torch.device("cuda", device)
        self.n_pages = PAGES_PER_GPU
        self.n_crops = PAGES_PER_GPU * CROPS_PER_PAGE
        self.n_tables = self.n_pages * TABLES_PER_PAGE if full else 0
        # ---- predictors (public API)
        if DET == "ppocrv4":
            det_task = predictors.OcrDetectionTask(model="db_pp", backbone="PPLCNetV3", state_dict=synth.pp_ocrv4_det_state_dict(0), device=device,
                                                   precision=precision)
        else:
            det_task = predictors.OcrDetectionTask(model="db_pp", state_dict=synth.dbnet_r18_state_dict(0), device=device, precision=precision)
        rec_task = predictors.OcrRecognitionTask(model="PP-OCRv4", state_dict=synth.pp_ocrv4_rec_state_dict(0, PP_REC_CLASSES), device=device,
                                                 vocab=[chr(33 + i) for i in range(PP_REC_CLASSES - 2)], precision=precision)
        lay_task = tsr_task = None
        if full:
            bb, nk, hd = synth.picodet_state_dicts(0, 5)
            lay_task = predictors.OcrLayoutTask(model="picodet", task_type="en", state_dict=(bb, nk, hd), device=device)
            sd = synth.lore_dla34_state_dict(0)
            sd["hm.2.bias"] = np.array(LORE_HM_BIAS, np.float32)
            tsr_task = predictors.OcrTableStructureTask(model="Lore", task_type="wtw", device=device, max_cells_per_image=1024,
                                                        state_dict=(sd, synth.lore_processor_state_dict(0)), precision=precision)
        self.system = OcrSystemTask(text_detector=det_task, text_recognizer=rec_task, table_structure_recognizer=tsr_task,
                                    layout_detector=lay_task)
        # ---- the same engine handles, driven directly for the device-resident leg
        self.det, self.rec, self.post = det_task.predictor, rec_task.predictor, rec_task.post
        self.layout = lay_task.predictor if full else None
        self.lore = tsr_task.predictor if full else None
        self.lore_proc = tsr_task.processor if full else None
        self.tsr_post = tsr_task.post if full else None
        self._tasks = [t for t in (det_task, rec_task, lay_task, tsr_task) if t is not None]
        # ---- inputs: host pages live in pinned memory and reach the API as a numpy array
        self.pages_pinned = torch.from_numpy(make_pages(rank, self.n_pages)).pin_memory()
        self.pages_np = self.pages_pinned.numpy()
        self.pages_dev = self.pages_pinned.to(dev)
        self.planted_maps = torch.from_numpy(make_prob_maps(rank, self.n_pages)).to(dev)
        self.src_hw = [(PAGE_H, PAGE_W)] * self.n_pages
        self.rec_crops = torch.empty((self.n_crops, 48, 1280, 3), dtype=torch.uint8, device=dev)
        self.crop_ws = (torch.empty((self.n_crops,), dtype=torch.int32, device=dev), torch.empty((self.n_crops, 2), dtype=torch.int32, device=dev),
                        torch.empty((self.n_crops, 3, 3), dtype=torch.float64, device=dev))
        self.prob_map = torch.empty((self.n_pages, 1, PAGE_H, PAGE_W), dtype=torch.float32, device=dev)
        self.flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)  # L2 flush buffer (> 126 MB)
        self._side = torch.cuda.Stream(device=dev)
        self._fork, self._join = torch.cuda.Event(), torch.cuda.Event()
        self.api_counts = {}
        if full:
            from pdf_table_b200 import predictors as P

            x0, y0, cw, ch = P.table_crop_rect(TABLE_BBOX, PAGE_H, PAGE_W)
            c, sc = np.array([cw / 2.0, ch / 2.0], dtype=np.float32), max(ch, cw) * 1.0
            meta = np.array([c[0], c[1], sc]).astype(np.int64)  # update_meta truncates the centre (processer_lore.py:112-130)
            self.table_rects = np.array([[p, x0, y0, cw, ch] for p in range(self.n_pages) for _ in range(TABLES_PER_PAGE)], np.int32)
            self.table_minv = np.stack([P.invert_affine(P.lore_affine(c, sc, 1024, 1024))] * self.n_tables)
            self.lore_inv = np.stack([P.lore_affine([np.float32(meta[0]), np.float32(meta[1])], np.float32(meta[2]), 256, 256, True)] * self.n_tables)
            self.lore_maps = torch.empty((self.n_tables, 256, 256, 24), dtype=torch.float32, device=dev)
            self.org_hw = [(PAGE_H, PAGE_W)] * self.n_pages
            self.layout_sf = [(800.0 / PAGE_H, 608.0 / PAGE_W)] * self.n_pages
            self.layout_tables = [[TABLE_BBOX] * TABLES_PER_PAGE for _ in range(self.n_pages)]

    @property
    def stages(self):
        s = ["det_preprocess_u8(fused)", "pp_ocrv4_det_forward(PPLCNetV3 + RSE-FPN + DBHead)" if DET == "ppocrv4" else "dbnet_r18_forward",
             "db_boxes(planted prob maps)",
             "crop_quads_for_rec(homography + warp + PP resize_norm_img width rule)", "pp_rec_preprocess_u8(fused)",
             "pp_ocrv4_rec_forward(PPLCNetV3 + SVTR + CTC head, per padded-width group)", "softmax+argmax+max", "ctc_greedy_decode"]
        if self.full:
            s = ["page_resize_u8(800x608)", "layout_preprocess_u8(fused)", "picodet_forward", "picodet_decode"] + s + [
                "crop_tables_for_tsr(slice + warpAffine)", "lore_preprocess_u8(fused)", "lore_dla34_dcn_forward", "lore_decode(wiz_rev)",
                "lore_cell_features", "lore_processor"]
        return s

    @property
    def engines(self):
        return [e for e in (self.det, self.rec, self.post, self.layout, self.lore, self.lore_proc) if e is not None] + \
               [t.post for t in self._tasks if getattr(t, "post", None) is not None and t.post is not self.post]

    # ---- device-resident leg (C-ABI level, inputs in HBM)
    def step_device(self):
        rec = {}
        if self.full and TWO_STREAMS:
            # the table-structure branch depends on the resident pages only: it runs on a second stream beside layout -> detect
            # -> recognise, so that the latency-bound kernels of one branch (contour tracing, crop warps, small attention) share the
            # SMs with the other's
            main = torch.cuda.current_stream()
            self._fork.record(main)
            with torch.cuda.stream(self._side):
                self._side.wait_event(self._fork)
                self._tables_branch(rec)
                self._join.record(self._side)
        if self.full:
            lay_in = self.post.resize_pages_u8(self.pages_dev, 608, 800)
            scores, dfl = self.layout.picodet_forward_u8(lay_in, flip=True)
            self.lay = self.post.picodet_decode(scores, dfl, self.org_hw, self.layout_sf, (800, 608))
        self.det.dbnet_forward_u8(self.pages_dev, MEAN, STD, 1.0 / 255.0, True, out=self.prob_map)
        boxes, counts = self.post.db_boxes(self.planted_maps, self.src_hw)
        from pdf_table_b200 import predictors

        crops, widths, sizes, _ = self.post.crop_boxes_for_rec(self.pages_dev, boxes, counts, CROPS_PER_PAGE, dst_h=48, dst_w_pad=1280,
                                                                out=self.rec_crops, ws=self.crop_ws, width_rule=1)
        r_ids, r_len, _ = predictors.pp_rec_launch_groups(self.rec, self.post, crops, widths, sizes)
        rec.update(boxes=boxes[:, :64].contiguous(), box_counts=counts, ids=r_ids, id_lens=r_len)
        if self.full and TWO_STREAMS:
            torch.cuda.current_stream().wait_event(self._join)
        elif self.full:
            self._tables_branch(rec)
        self.record = rec
        return rec

    def _tables_branch(self, rec):
        tables = self.tsr_post.crop_tables_for_tsr(self.pages_dev, self.table_rects, self.table_minv, 1024, 1024)
        self.lore.lore_detect_forward_u8(tables, out=self.lore_maps)
        dec = self.tsr_post.lore_decode(self.lore_maps, None, None, None, self.lore_inv)
        feat, offsets = self.lore.lore_cell_features(dec, max_rows=self.n_tables * 1024)
        logi = self.lore_proc.lore_process_forward(feat, offsets)[1]
        idx = offsets[:-1].long()[:, None] + torch.arange(256, device=logi.device)[None, :]
        rec.update(cells=dec["polygons"][:, :256].contiguous(), cell_counts=dec["counts"],
                   cell_logi=logi[idx.clamp_(max=int(logi.shape[0]) - 1)])

    # ---- end-to-end leg (public API: numpy pages in, Python results out)
    def step_e2e(self):
        planted = self.planted_maps
        out = self.system.predict_pages(self.pages_np, layout_tables=self.layout_tables if self.full else None,
                                        det_kwargs={"prob_override": lambda prob, idx: planted if len(idx) == planted.shape[0] else planted[idx]},
                                        keep_device_record=True)
        torch.cuda.current_stream().synchronize()
        self.api_out = out
        return self.system.device_record

    def stream_e2e(self, steps: int):
        """`steps` e2e steps through OcrSystemTask.predict_stream: every step's pages go up from pinned host memory inside the
        loop (the copy of step i + 1 runs beside the device work of step i), every step's results come back as Python objects."""
        planted = self.planted_maps
        for out in self.system.predict_stream((self.pages_np for _ in range(steps)), layout_tables=self.layout_tables if self.full else None,
                                              det_kwargs={"prob_override": lambda prob, idx: planted if len(idx) == planted.shape[0] else planted[idx]},
                                              keep_device_record=True):
            self.api_out = out
            yield self.system.device_record

    def e2e_bytes(self):
        """(h2d, d2h) bytes of ONE e2e step, counted by the predictors' own transfer helpers (predictors.TRANSFER)."""
        from pdf_table_b200 import predictors

        predictors.TRANSFER["h2d"] = predictors.TRANSFER["d2h"] = 0
        self.step_e2e()
        return int(predictors.TRANSFER["h2d"]), int(predictors.TRANSFER["d2h"])

    def flush_l2(self):
        self.flush.fill_(1)


def gather_record(dist, rec, world: int, n_pages: int, n_tables: int):
    """The ONE collective of the path: all-gather of the packed decoded results of this batch."""
    from pdf_table_b200 import sharding

    # crop counts differ per rank in the API leg (detected boxes): the pad size is the fixed slot count of the workload
    cap = max(int(rec["ids"].shape[0]), PAGES_PER_GPU * CROPS_PER_PAGE)
    return sharding.all_gather_results(rec, [n_pages] * world, [cap] * world, table_sizes=[n_tables] * world if n_tables else ())


def aggregate(recs):
    agg = {}
    for r in recs:
        k = agg.setdefault(r["kernel"], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "n": 0})
        k["ms"] += r["ms"]
        k["flops"] += r["flops"]
        k["bytes"] += r["bytes"]
        k["n"] += 1
    return agg


def roofline_of(agg, peaks, peak_src, workload_key):
    tot_ms = sum(k["ms"] for k in agg.values())
    top = max(agg, key=lambda k: agg[k]["ms"])
    tk = agg[top]
    traffic, traffic_src = measured_traffic(top, workload_key)
    common = {"kernel": top, "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": tk["bytes"] / tk["n"],
              "algorithmic_flops_per_launch": tk["flops"] / tk["n"], "launches": tk["n"], "avg_launch_ms": tk["ms"] / tk["n"],
              "share_of_step": tk["ms"] / tot_ms, "hbm_achieved_gbs": tk["bytes"] / (tk["ms"] / 1e3) / 1e9}
    # a kernel whose arithmetic intensity is below the ridge (peak FLOP/s / peak B/s) is judged against the HBM roofline
    ridge = peaks["bf16_tflops_sustained"] * 1e12 / (peaks["hbm_gbs"] * 1e9)
    if tk["flops"] > 0 and tk["flops"] / max(tk["bytes"], 1.0) >= ridge * 0.25:
        ach = tk["flops"] / (tk["ms"] / 1e3) / 1e12
        return {"bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops_sustained"], "unit": "TFLOP/s",
                "frac": ach / peaks["bf16_tflops_sustained"], "peak_source": peak_src + ", sustained bf16 (kernel timed inside a long step)", **common}
    ach = tk["bytes"] / (tk["ms"] / 1e3) / 1e9
    return {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "peak_source": peak_src, **common}


def kernel_table(agg, steps):
    return {k: {"ms_per_step": v["ms"] / steps, "launches_per_step": v["n"] / steps,
                "tflops": (v["flops"] / (v["ms"] / 1e3) / 1e12) if v["flops"] else None,
                "gbs": v["bytes"] / (v["ms"] / 1e3) / 1e9} for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}


def timed_steps(step, flush, steps, barrier):
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    barrier()
    for a, b in evs:
        flush()
        a.record()
        step()
        b.record()
    barrier()
    return sum(a.elapsed_time(b) for a, b in evs)


def timed_loop(step, steps, barrier):
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        step()
    b.record()
    barrier()
    return a.elapsed_time(b)


def max_over_ranks(dist, vals, dev):
    t = torch.tensor(vals, dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t]


# --------------------------------------------------------------------------------------- secondary blocks
def block_rec_sweep(wl: Cascade, rank, world, dist, args, barrier, peaks, peak_src):
    """BASELINE configs[3]: ConvNextViT on 4096 text-line crops 32x320, dealt round-robin over the N GPUs (strong scaling),
    one all-gather of the decoded ids per batch inside the timed step."""
    from pdf_table_b200 import sharding

    dev = torch.device("cuda", wl.device)
    counts = [len(range(r, SWEEP_CROPS, world)) for r in range(world)]
    n = counts[rank]
    from pdf_table_b200 import predictors, synth

    task = predictors.OcrRecognitionTask(model="ConvNextViT", state_dict=synth.convnext_vit_state_dict(0), device=wl.device)
    n_labels = int(task.predictor._lib.dv_convnextvit_labels(task.predictor._h))
    task.label_mapping = {i + 2: ch for i, ch in enumerate(synthetic_vocab(n_labels))}
    rec, post = task.predictor, task.post
    crops_np = make_sweep_crops(n, rank, world)
    crops_list = list(crops_np)
    crops_dev = torch.from_numpy(crops_np).to(dev)
    ids = torch.empty((n, 201), dtype=torch.int32, device=dev)
    none = {"boxes": torch.zeros((0, 1, 8), dtype=torch.float32, device=dev), "box_counts": torch.zeros((0,), dtype=torch.int32, device=dev)}

    def gather(out, ln):
        if dist is not None:
            sharding.all_gather_results({**none, "ids": out, "id_lens": ln}, [0] * world, counts)

    def step_device():
        rec.convnextvit_forward_u8(crops_dev, ids=ids)
        out, ln, _ = post.ctc_collapse(ids)
        gather(out, ln)

    def step_e2e():  # the public call: a list of numpy crops in, strings out (keepratio_resize + padding on the host as the reference)
        texts = task(crops_list)
        if dist is not None:  # the decoded ids of the last chunk stand for the batch's exchange (same bytes: all chunks are equal)
            gather(ids_out[0], ids_out[1])
        return texts

    ids_out = post.ctc_collapse(ids)[:2]

    l0 = rec.launch_count + post.launch_count
    step_device()
    launches = rec.launch_count + post.launch_count - l0
    for _ in range(max(args.warmup, 3) - 1):
        step_device()
    dev_ms = timed_steps(step_device, wl.flush_l2, args.steps, barrier)
    step_e2e()
    e2e_ms = timed_loop(step_e2e, args.steps, barrier)
    rec.profile_begin()
    for _ in range(args.steps):
        wl.flush_l2()
        step_device()
    agg = aggregate(rec.profile_report())
    dev_ms, e2e_ms = max_over_ranks(dist, [dev_ms, e2e_ms], dev)
    total = SWEEP_CROPS * args.steps
    return {"metric": "text_line_crops_per_sec", "value": total / (dev_ms / 1e3), "unit": "crops/s", "ms_per_step": dev_ms / args.steps,
            "scaling": "strong", "higher_is_better": True, "dtype": "f16",
            "config": {"workload": f"BASELINE configs[3]: ConvNextViT recogniser, {SWEEP_CROPS} synthetic text-line crops {SWEEP_H}x{SWEEP_W} (3 chunks "
                                   f"each) dealt round-robin over {world} GPU(s)", "crops_per_gpu": counts,
                       "stages": ["rec_preprocess_u8(fused)", "convnextvit_forward+argmax", "ctc_collapse"] + (["all_gather(ids)"] if world > 1 else []),
                       "model_gflop_per_crop": rec.model_flops / max(n, 1) / 1e9},
            "e2e": {"value": total / (e2e_ms / 1e3), "unit": "crops/s", "h2d_bytes_per_step": int(crops_np.nbytes) * world,
                    "d2h_bytes_per_step": int(n * 201 * 4 + n * 4) * world, "ms_per_step": e2e_ms / args.steps,
                    "api": "OcrRecognitionTask.__call__(list of numpy crops) -> list[str]"},
            "roofline": roofline_of(agg, peaks, peak_src, "rec_sweep"), "gpu_launches": int(launches * args.steps),
            "kernels": kernel_table(agg, args.steps)}


def block_pp_rec(wl: Cascade, rank, world, dist, args, barrier, peaks, peak_src):
    """BASELINE metric, first quantity on its own model: text-line crops/s through the PP-OCRv4 recogniser (SVTR-LCNet: a4 -> a5
    -> a6) on 4096 synthetic crops at the nominal 48x320 shape, dealt round-robin over the N GPUs (strong scaling), one
    all-gather of the decoded ids per batch inside the timed step."""
    from pdf_table_b200 import predictors, sharding, synth

    dev = torch.device("cuda", wl.device)
    counts = [len(range(r, SWEEP_CROPS, world)) for r in range(world)]
    n = counts[rank]
    vocab = [chr(33 + i) for i in range(PP_REC_CLASSES - 2)]
    task = predictors.OcrRecognitionTask(model="PP-OCRv4", state_dict=synth.pp_ocrv4_rec_state_dict(0, PP_REC_CLASSES), vocab=vocab, device=wl.device)
    rec, post = task.predictor, task.post
    import cv2

    crops_np = np.stack([cv2.resize(c, (320, 48)) for c in make_sweep_crops(n, rank, world)])  # 48 x 320: ratio 6.67 = the nominal shape
    crops_list = list(crops_np)
    crops_dev = torch.from_numpy(crops_np).to(dev)
    widths_dev = torch.full((n,), 320, dtype=torch.int32, device=dev)
    none = {"boxes": torch.zeros((0, 1, 8), dtype=torch.float32, device=dev), "box_counts": torch.zeros((0,), dtype=torch.int32, device=dev)}
    t_steps = rec.rec_time_steps(48, 320)

    def gather(out, ln):
        if dist is not None:
            sharding.all_gather_results({**none, "ids": out, "id_lens": ln}, [0] * world, counts)

    def step_device():
        ids, maxp = rec.rec_forward_u8(crops_dev, widths_dev)
        out, ln, _ = post.ctc_collapse(ids, maxp)
        gather(out, ln)
        return out, ln

    last = [None]

    def step_e2e():  # the public call: a list of numpy crops in, strings out (aspect sort, batch plan and cv2.resize on the host)
        texts = task(crops_list)
        if dist is not None:
            gather(*last[0])
        return texts

    engines = (rec, post)
    l0 = sum(e.launch_count for e in engines)
    last[0] = step_device()
    launches = sum(e.launch_count for e in engines) - l0
    for _ in range(max(args.warmup, 3) - 1):
        step_device()
    dev_ms = timed_steps(step_device, wl.flush_l2, args.steps, barrier)
    step_e2e()
    e2e_steps = max(1, min(args.steps, 5))
    e2e_ms = timed_loop(step_e2e, e2e_steps, barrier) * args.steps / e2e_steps
    for e in engines:
        e.profile_begin()
    for _ in range(args.steps):
        wl.flush_l2()
        step_device()
    recs = []
    for e in engines:
        recs += e.profile_report()
    agg = aggregate(recs)
    dev_ms, e2e_ms = max_over_ranks(dist, [dev_ms, e2e_ms], dev)
    total = SWEEP_CROPS * args.steps
    return {"metric": "text_line_crops_per_sec", "value": total / (dev_ms / 1e3), "unit": "crops/s", "ms_per_step": dev_ms / args.steps,
            "scaling": "strong", "higher_is_better": True, "dtype": "f16",
            "config": {"workload": f"BASELINE metric (PP-OCRv4 rec): SVTR-LCNet recogniser, {SWEEP_CROPS} synthetic text-line crops 48x320 dealt "
                                   f"round-robin over {world} GPU(s); T = {t_steps} steps, {PP_REC_CLASSES} classes (en dictionary)",
                       "rec_model": "PP-OCRv4 rec = PPLCNetV3-0.95 + SVTR neck + CTC head, the published architecture the hub ONNX was exported from "
                                    "(the ONNX file itself is not in the reference tree), seeded random weights",
                       "crops_per_gpu": counts, "stages": ["rec_preprocess_u8(fused: /255, -0.5, /0.5, zero pad)", "pplcnetv3_svtr_forward",
                                                           "softmax+argmax+max", "ctc_greedy_decode"] + (["all_gather(ids)"] if world > 1 else []),
                       "model_gflop_per_crop": rec.model_flops / max(n, 1) / 1e9},
            "e2e": {"value": total / (e2e_ms / 1e3), "unit": "crops/s", "h2d_bytes_per_step": int(crops_np.nbytes + 4 * n) * world,
                    "d2h_bytes_per_step": int(n * t_steps * 4 + n * 8) * world, "ms_per_step": e2e_ms / args.steps,
                    "api": "OcrRecognitionTask(model='PP-OCRv4').__call__(list of numpy crops) -> list[str]"},
            "roofline": roofline_of(agg, peaks, peak_src, "pp_rec"), "gpu_launches": int(launches * args.steps),
            "kernels": kernel_table(agg, args.steps)}


def block_lore(wl: Cascade, rank, world, dist, args, barrier, peaks, peak_src):
    """BASELINE configs[2]: Lore (DLA-34 + DCNv2, wtw), 16 x 1024x1024 table crops per GPU: detect + decode + cell features +
    processor.  e2e through OcrTableStructureTask.__call__ on numpy crops (host cv2.warpAffine pre-process included)."""
    dev = torch.device("cuda", wl.device)
    imgs_np, inv = make_table_crops(rank, LORE_BATCH)
    imgs_dev = torch.from_numpy(imgs_np).to(dev)
    maps = torch.empty((LORE_BATCH, 256, 256, 24), dtype=torch.float32, device=dev)
    engines = (wl.lore, wl.lore_proc, wl.post)
    task = wl.system.table_structure_recognizer
    crops_list = list(imgs_np)

    def step_device():
        wl.lore.lore_detect_forward_u8(imgs_dev, out=maps)
        dec = wl.post.lore_decode(maps, None, None, None, inv)
        feat, offsets = wl.lore.lore_cell_features(dec, max_rows=LORE_BATCH * 1024)
        return dec, wl.lore_proc.lore_process_forward(feat, offsets)

    def step_e2e():
        return task(crops_list)

    l0 = sum(e.launch_count for e in engines)
    dec, _ = step_device()
    launches = sum(e.launch_count for e in engines) - l0
    for _ in range(max(args.warmup, 3) - 1):
        step_device()
    cells = int(dec["counts"].sum())
    dev_ms = timed_steps(step_device, wl.flush_l2, args.steps, barrier)
    step_e2e()
    e2e_ms = timed_loop(step_e2e, args.steps, barrier)
    for e in engines:
        e.profile_begin()
    for _ in range(args.steps):
        wl.flush_l2()
        step_device()
    recs = []
    for e in engines:
        recs += e.profile_report()
    agg = aggregate(recs)
    dev_ms, e2e_ms = max_over_ranks(dist, [dev_ms, e2e_ms], dev)
    total = LORE_BATCH * world * args.steps
    return {"metric": "table_images_per_sec", "value": total / (dev_ms / 1e3), "unit": "images/s", "ms_per_step": dev_ms / args.steps,
            "scaling": "weak", "higher_is_better": True, "dtype": "f16",
            "config": {"workload": f"BASELINE configs[2]: Lore DLA-34 + DCNv2 (wtw) table structure, {LORE_BATCH} synthetic 1024x1024 table crops per GPU",
                       "stages": ["lore_preprocess_u8(fused)", "lore_dla34_dcn_forward", "lore_decode(wiz_rev)", "lore_cell_features", "lore_processor"],
                       "cells_per_step": cells, "model_gflop_per_image": wl.lore.model_flops / LORE_BATCH / 1e9},
            "e2e": {"value": total / (e2e_ms / 1e3), "unit": "images/s", "h2d_bytes_per_step": int(imgs_np.nbytes), "ms_per_step": e2e_ms / args.steps,
                    "d2h_bytes_per_step": int(LORE_BATCH * 8 + 4 + cells * (8 + 4) * 4),
                    "api": "OcrTableStructureTask.__call__ on a list of numpy crops (cv2.warpAffine on the host as the reference, dicts out)"},
            "roofline": roofline_of(agg, peaks, peak_src, "lore"), "gpu_launches": int(launches * args.steps), "kernels": kernel_table(agg, args.steps)}


def block_lore_wireless(wl: Cascade, rank, local_rank, world, dist, args, barrier, peaks, peak_src):
    """Lore `wireless` (ResNet-18 key-point detector, 768 x 768 upper-left frame, 2-D position embeddings), 16 synthetic table
    crops per GPU: detect + decode + cell features + position embeddings + processor; e2e through OcrTableStructureTask.__call__."""
    from pdf_table_b200 import predictors, synth

    dev = torch.device("cuda", local_rank)
    crops_list = [synth.synthetic_page(rank * 1000 + 700 + i % 4, 900, 768) for i in range(LORE_BATCH)]
    pre = [predictors.lore_preprocess(c, (768, 768), upper_left=True) for c in crops_list[:4]]
    imgs_dev = torch.from_numpy(np.stack([pre[i % 4][0] for i in range(LORE_BATCH)])).to(dev)
    sd = synth.lore_resnet18_state_dict(0)
    # seeded random weights: calibrate the heat map's bias so that ~150 cells per table pass the 0.2 gate (a trained detector's
    # count; unshifted, the random heat map saturates the decode's 3000-cell cap and the step would time the processor only)
    probe = predictors.OcrTableStructureTask(model="Lore", task_type="wireless", device=local_rank, state_dict=(sd, synth.lore_processor_state_dict(0)))
    hm = probe.predictor.lore_detect_forward_u8(imgs_dev[:4])[..., 0].clamp(1e-6, 1 - 1e-6)
    peaks_at = torch.nn.functional.max_pool2d(hm[:, None], 3, 1, 1)[:, 0] == hm
    logit = torch.log(hm / (1 - hm))[peaks_at]
    kth = float(torch.topk(logit, min(150 * 4, logit.numel())).values[-1])
    for e in (probe.predictor, probe.processor, probe.post):
        e.close()
    sd["hm.8.bias"] = sd["hm.8.bias"] + np.float32(np.log(0.2 / 0.8) - kth)
    task = predictors.OcrTableStructureTask(model="Lore", task_type="wireless", device=local_rank, host_warp=True,
                                            state_dict=(sd, synth.lore_processor_state_dict(0)))
    inv = np.stack([predictors.lore_affine_upper_left([np.float32(0), np.float32(0)], np.float32(pre[i % 4][1][2]), 192, 192, True) for i in range(LORE_BATCH)])
    maps = torch.empty((LORE_BATCH, 192, 192, 24), dtype=torch.float32, device=dev)
    det, proc, post = task.predictor, task.processor, task.post
    engines = (det, proc, post)

    def step_device():
        det.lore_detect_forward_u8(imgs_dev, out=maps)
        dec = post.lore_decode(maps, None, None, None, inv, wiz_rev=False, vis_thresh=0.2)
        feat, offsets = det.lore_cell_features(dec, max_rows=LORE_BATCH * 1024)
        proc.lore_add_position_embeddings(feat, dec, offsets)
        return dec, proc.lore_process_forward(feat, offsets)

    l0 = sum(e.launch_count for e in engines)
    dec, _ = step_device()
    launches = sum(e.launch_count for e in engines) - l0
    for _ in range(max(args.warmup, 3) - 1):
        step_device()
    cells = int(dec["counts"].sum())
    dev_ms = timed_steps(step_device, wl.flush_l2, args.steps, barrier)
    task(crops_list)
    e2e_ms = timed_loop(lambda: task(crops_list), args.steps, barrier)
    for e in engines:
        e.profile_begin()
    for _ in range(args.steps):
        wl.flush_l2()
        step_device()
    recs = []
    for e in engines:
        recs += e.profile_report()
    agg = aggregate(recs)
    dev_ms, e2e_ms = max_over_ranks(dist, [dev_ms, e2e_ms], dev)
    total = LORE_BATCH * world * args.steps
    out = {"metric": "table_images_per_sec", "value": total / (dev_ms / 1e3), "unit": "images/s", "ms_per_step": dev_ms / args.steps,
           "scaling": "weak", "higher_is_better": True, "dtype": "f16",
           "config": {"workload": f"Lore ResNet-18 (wireless) table structure, {LORE_BATCH} synthetic table crops per GPU warped to 768x768",
                      "cells_per_step": cells, "model_gflop_per_image": det.model_flops / LORE_BATCH / 1e9},
           "e2e": {"value": total / (e2e_ms / 1e3), "unit": "images/s", "ms_per_step": e2e_ms / args.steps,
                   "h2d_bytes_per_step": int(LORE_BATCH * 768 * 768 * 3), "d2h_bytes_per_step": int(LORE_BATCH * 8 + 4 + cells * (8 + 4) * 4),
                   "api": "OcrTableStructureTask(task_type='wireless').__call__ on a list of numpy crops (cv2.warpAffine on the host, dicts out)"},
           "roofline": roofline_of(agg, peaks, peak_src, "lore_wireless"), "gpu_launches": int(launches * args.steps), "kernels": kernel_table(agg, args.steps)}
    for e in engines:
        e.close()
    return out


def block_fp32x(rank, local_rank, world, dist, args, barrier):
    """The same cascade step with precision="fp32x" in the three networks the north star's 1e-3 bound addresses (detector,
    recogniser, Lore detector): the price of the precise mode, device-resident leg only."""
    wl = Cascade(rank, local_rank, FULL, precision="fp32x")
    for _ in range(2):
        wl.step_device()
    torch.cuda.synchronize()
    steps = max(2, min(args.steps, 3))
    ms = timed_steps(wl.step_device, wl.flush_l2, steps, barrier)
    ms = max_over_ranks(dist, [ms], torch.device("cuda", local_rank))[0]
    for t in wl._tasks:
        for e in (getattr(t, "predictor", None), getattr(t, "processor", None), getattr(t, "post", None)):
            if e is not None:
                e.close()
    del wl
    torch.cuda.empty_cache()
    return {"metric": "pages_per_sec", "value": PAGES_PER_GPU * world * steps / (ms / 1e3), "unit": "pages/s", "ms_per_step": ms / steps, "steps": steps,
            "dtype": "f16 split pairs (hi + lo) / fp32 accumulate", "scaling": "weak",
            "config": {"workload": "the default cascade with precision='fp32x' in the PP-OCRv4 detector, the PP-OCRv4 recogniser and the Lore detector "
                                   "(outputs within 1e-3 of the fp32 oracle, DESIGN.md 5); PicoDet and the Lore processor as in the default line"}}


# --------------------------------------------------------------------------------------- CPU arm
def cpu_reference_step(sample_pages: np.ndarray, sd, rec_sd, sample_maps):
    """The reference's algorithm for the same stages on the host cores (oracle/ restatement of PPOcrDetectionPreprocessor +
    DBModel + DBPostProcess, crop_image + PPOcrRecPreProcessor (one crop per call, as the orchestrator does) + the PP-OCRv4
    recogniser + CTCLabelDecode)."""
    import cv2
    import math

    from oracle import crop_ref, ctc_ref, db_post_ref, dbnet_ref, pp_det_ref, pp_rec_ref

    mean = np.array(MEAN, np.float32).reshape(1, 1, 3)
    std = np.array(STD, np.float32).reshape(1, 1, 3)
    for pg in sample_pages:
        img = pg[:, :, ::-1].astype("float32") * np.float32(1.0 / 255.0)
        img = (img - mean) / std
        x = torch.from_numpy(np.ascontiguousarray(img.transpose(2, 0, 1))[None])
        if DET == "ppocrv4":
            pp_det_ref.pp_det_forward(sd, x)
        else:
            dbnet_ref.dbnet_r18_forward(sd, x)
    for pg, m in zip(sample_pages, sample_maps):
        boxes = db_post_ref.db_postprocess(m[0], np.array([PAGE_H, PAGE_W, 1.0, 1.0]), (PAGE_H, PAGE_W, 3))
        for b in boxes[:CROPS_PER_PAGE]:  # OcrCommonUtils.crop_image per detected quad, one recogniser call per crop (ocr_system_task.py:300-313)
            try:
                crop = crop_ref.crop_image(pg, b.reshape(4, 2))
            except Exception:  # an empty crop: cv2 raises, the reference's orchestrator skips the box
                continue
            h, w = crop.shape[:2]
            img_w = max(min(int(48 * max(w * 1.0 / h, 320 / 48)), 1280), 16)  # resize_norm_img for a batch of one
            rw = min(img_w, max(math.ceil(48 * (w / float(h))), 16))
            x = np.zeros((1, 3, 48, img_w), np.float32)
            x[0, :, :, :rw] = (cv2.resize(crop, (rw, 48)).astype("float32").transpose(2, 0, 1) / 255 - 0.5) / 0.5
            ctc_ref.ctc_greedy_ids(pp_rec_ref.pp_rec_forward(rec_sd, torch.from_numpy(x)).numpy())
    if FULL:
        cpu_full_extra(len(sample_pages))


_FULL_CPU = {}


def cpu_full_extra(n_pages: int):
    """The reference's algorithm for the layout and table-structure stages on the host cores (oracle/ restatements of
    LCNet + CSP-PAN + PicoHead + OCRPicodetPostProcessor and of get_dla_dcn + process_detect_output + LoreProcessModel)."""
    from oracle import picodet_net_ref, picodet_ref
    from pdf_table_b200 import synth

    if not _FULL_CPU:
        _FULL_CPU["pico"] = synth.picodet_state_dicts(0, 5)
        _FULL_CPU["pages"] = make_layout_pages(make_pages(0, 4))
    mean = np.array(MEAN, np.float32).reshape(1, 1, 3)
    std = np.array(STD, np.float32).reshape(1, 1, 3)
    for i in range(n_pages):
        pg = _FULL_CPU["pages"][i % 4]
        x = ((pg[:, :, ::-1].astype("float32") * np.float32(1.0 / 255.0) - mean) / std).transpose(2, 0, 1)[None]
        s, d = picodet_net_ref.picodet_forward(*_FULL_CPU["pico"], torch.from_numpy(np.ascontiguousarray(x)), 5)
        picodet_ref.picodet_decode([t.numpy() for t in s], [t.numpy() for t in d], [PAGE_H, PAGE_W], [800.0 / PAGE_H, 608.0 / PAGE_W], [800, 608])
    cpu_lore(n_pages)


def cpu_lore(n_images: int):
    from oracle import lore_decode_ref, lore_net_ref, lore_processor_ref
    from pdf_table_b200 import synth

    if "lore" not in _FULL_CPU:
        sd = synth.lore_dla34_state_dict(0)
        sd["hm.2.bias"] = np.array(LORE_HM_BIAS, np.float32)
        _FULL_CPU["lore"] = sd
        _FULL_CPU["proc"] = synth.lore_processor_state_dict(0)
        _FULL_CPU["tables"] = make_table_crops(0, 4)
    lmean = np.array([0.408, 0.447, 0.470], np.float32).reshape(1, 1, 3)
    lstd = np.array([0.289, 0.274, 0.278], np.float32).reshape(1, 1, 3)
    for i in range(n_images):
        tb = _FULL_CPU["tables"][0][i % 4]
        x = ((tb / 255. - lmean) / lstd).astype(np.float32).transpose(2, 0, 1)[None]
        out = lore_net_ref.lore_dla34_forward(_FULL_CPU["lore"], torch.from_numpy(np.ascontiguousarray(x)))
        meta = np.array([512, 512, 1024, 1024, 1024, 256, 256])
        dec = lore_decode_ref.lore_decode(torch.sigmoid(out["hm"])[0].numpy(), out["reg"][0].numpy(), out["wh"][0].numpy(), out["st"][0].numpy(),
                                          out["ax"][0].numpy(), out["cr"][0].numpy(), meta)
        if len(dec["logi_feat"]):
            lore_processor_ref.lore_processor_forward(_FULL_CPU["proc"], torch.from_numpy(dec["logi_feat"]))


def cpu_rec_sweep(n: int):
    from oracle import convnextvit_ref
    from pdf_table_b200 import synth

    if "rec" not in _FULL_CPU:
        _FULL_CPU["rec"] = {k: torch.from_numpy(v) for k, v in synth.convnext_vit_state_dict(0).items()}
    crops = list(make_sweep_crops(n))
    for i in range(0, n, 16):
        convnextvit_ref.greedy_ids(convnextvit_ref.convnextvit_forward(_FULL_CPU["rec"], convnextvit_ref.preprocess(crops[i:i + 16])))


def host_threads() -> int:
    """All host cores for the CPU arm, also under torchrun (which exports OMP_NUM_THREADS=1 to every rank)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    return cores


class CpuArm:
    def __init__(self, n_pages: int):
        from pdf_table_b200 import synth

        self.n_pages = n_pages
        self.sd = {k: torch.from_numpy(v) for k, v in (synth.pp_ocrv4_det_state_dict(0) if DET == "ppocrv4" else synth.dbnet_r18_state_dict(0)).items()}
        self.rec_sd = {k: torch.from_numpy(v) for k, v in synth.pp_ocrv4_rec_state_dict(0, PP_REC_CLASSES).items()}
        self.pages = make_pages(0, n_pages)
        self.maps = make_prob_maps(0, n_pages)

    def step(self):
        cpu_reference_step(self.pages, self.sd, self.rec_sd, self.maps)

    def sample(self):
        return (f"{self.n_pages} of {PAGES_PER_GPU} pages 960x960 per step, each with {CROPS_PER_PAGE} text-line crop slots cut from the page at the "
                "detected quads (crop_image + resize_norm_img, one crop per recogniser call as the reference's orchestrator) through the PP-OCRv4 "
                "recogniser + CTC decode" +
                (", PicoDet layout and Lore table structure on one table crop per page" if FULL else "") + ", oracle/ restatement in torch fp32 on the host cores")


def cpu_pp_rec(n: int):
    import cv2

    from oracle import ctc_ref, pp_rec_ref
    from pdf_table_b200 import synth

    if "pp_rec" not in _FULL_CPU:
        _FULL_CPU["pp_rec"] = {k: torch.from_numpy(v) for k, v in synth.pp_ocrv4_rec_state_dict(0, PP_REC_CLASSES).items()}
    crops = np.stack([cv2.resize(c, (320, 48)) for c in make_sweep_crops(n)])
    for i in range(0, n, 6):  # batches of six, as PPOcrRecPreProcessor forms them
        x = ((crops[i:i + 6].astype("float32").transpose(0, 3, 1, 2) / 255 - 0.5) / 0.5).astype(np.float32)
        ctc_ref.ctc_greedy_ids(pp_rec_ref.pp_rec_forward(_FULL_CPU["pp_rec"], torch.from_numpy(x)).numpy())


def cpu_blocks():
    """Bounded CPU samples of the secondary workloads (once each, not per step)."""
    out = {}
    cpu_pp_rec(6)
    t0 = time.perf_counter()
    cpu_pp_rec(96)
    dt = time.perf_counter() - t0
    out["pp_rec"] = {"value": 96 / dt, "unit": "crops/s", "kind": "port", "sample": f"96 of {SWEEP_CROPS} crops 48x320 once in batches of 6, {dt:.1f} s"}
    t0 = time.perf_counter()
    cpu_rec_sweep(32)
    dt = time.perf_counter() - t0
    out["rec_sweep"] = {"value": 32 / dt, "unit": "crops/s", "kind": "port", "sample": f"32 of {SWEEP_CROPS} crops once, {dt:.1f} s"}
    if FULL:
        t0 = time.perf_counter()
        cpu_lore(1)
        dt = time.perf_counter() - t0
        out["lore"] = {"value": 1 / dt, "unit": "images/s", "kind": "port", "sample": f"1 of {LORE_BATCH} table crops once (detector + decode + processor), {dt:.1f} s"}
    return out


def run_reference(args, rank: int):
    if rank != 0:
        return
    cores = host_threads()
    arm = CpuArm(1 if FULL else 2)
    for _ in range(max(0, min(args.warmup, 1))):
        arm.step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        arm.step()
    dt = (time.perf_counter() - t0) / args.steps
    v = arm.n_pages / dt
    blocks = cpu_blocks()
    for b in blocks.values():
        b["cores"] = cores
    line = {
        "impl": "reference", "metric": "pages_per_sec", "value": v, "unit": "pages/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(),
        "cpu_baseline": {"value": v, "unit": "pages/s", "cores": cores, "kind": "port", "sample": arm.sample()},
        "e2e": {"value": v, "unit": "pages/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "blocks": blocks,
    }
    print(json.dumps(line), flush=True)


def workload_config():
    cfg = {
        "workload": "BASELINE configs[1]: DB detect + text-line recognise, batch=32 synthetic pages 960x960 per GPU",
        "det_model": ("PP-OCRv4 det = PPLCNetV3-0.75 + RSE-FPN + DBHead (the published architecture the hub ONNX was exported from, SURVEY.md a2), "
                      "seeded random weights" if DET == "ppocrv4" else
                      "DBNet-R18 (the reference's in-tree model='db' network, --det dbnet_r18), seeded random weights"),
        "rec_model": f"PP-OCRv4 rec = PPLCNetV3-0.95 + SVTR neck + CTC head (the published architecture the hub ONNX was exported from, SURVEY.md "
                     f"a5; {PP_REC_CLASSES} classes), {CROPS_PER_PAGE} crops per page cut on the device from the db_boxes quads of that page "
                     "(crop_image + resize_norm_img: 48 high, each crop padded to its own width as the reference's one-crop-per-call flow does; "
                     "crops of equal padded width share a launch), CTC greedy decode of the network's own output, seeded random weights",
        "db_post_stage": "db_boxes on planted probability maps (40 analytic text-line blobs per page, every one a box; + specks): with random weights the "
                         "detector's own map is texture noise; the network still runs and its map is discarded",
        "pages_per_gpu": PAGES_PER_GPU, "page": [PAGE_H, PAGE_W, 3],
        "l2": "flushed between timed steps (256 MiB write); activations per step exceed L2",
        "streams": "2: the table-structure branch (crop + Lore + decode + processor) runs on its own CUDA stream beside layout -> detect -> recognise, "
                   "in the device leg and in OcrSystemTask.predict_pages alike; the per-kernel table / roofline come from a one-stream pass of the same step" if TWO_STREAMS else "1",
        "parallelism": "page-sharded replicas, one process per GPU; one all-gather of the packed decoded results per batch inside the timed step",
    }
    if FULL:
        cfg["workload"] = ("BASELINE configs[4] per GPU: full cascade PicoDet layout -> DB detect -> text-line recognise -> Lore table structure, "
                           "32 synthetic pages 960x960 per GPU, one 1024x1024 table crop per page")
        cfg["layout_model"] = "PicoDet LCNet-x1.0 + CSP-PAN + PicoHead on 800x608 (in-tree modules, seeded random weights), page resized on the device"
        cfg["tsr_model"] = ("Lore DLA-34 + DCNv2 (wtw) + processor, one table per page cut from the resident page at a fixed layout box and warped into "
                            "the network frame on the device (dv_crop_tables_for_tsr: crop_image_by_box + cv2.warpAffine, bit-exact vs cv2), heat-map "
                            "bias shifted so ~100 cells per table are selected")
    return cfg


# --------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-blocks", action="store_true", help="skip the configs[3] / configs[2] blocks (profiling runs)")
    ap.add_argument("--cascade", default="full", choices=["ocr", "full"],
                    help="full = BASELINE configs[4] per GPU (the default line); ocr = configs[1] (DB detect + recognise only)")
    ap.add_argument("--det", default="ppocrv4", choices=["ppocrv4", "dbnet_r18"],
                    help="detector network of the cascade: the PP-OCRv4 det graph (default, BASELINE configs[1] / [4]) or the in-tree DBNet-R18")
    args = ap.parse_args()
    global FULL, DET, TWO_STREAMS
    FULL = args.cascade == "full"
    DET = args.det
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return
    args.warmup = max(args.warmup, 3)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    wl = Cascade(rank, local_rank, FULL)

    def step_device():
        rec = wl.step_device()
        if dist is not None:
            gather_record(dist, rec, world, wl.n_pages, wl.n_tables)

    def step_e2e():
        rec = wl.step_e2e()
        if dist is not None:
            gather_record(dist, rec, world, wl.n_pages, wl.n_tables)

    l0 = sum(e.launch_count for e in wl.engines)
    step_device()
    torch.cuda.synchronize()
    launches_per_step = sum(e.launch_count for e in wl.engines) - l0

    # ---- device-resident timing: K steps, each bracketed by events, L2 flushed in between (untimed)
    for _ in range(args.warmup - 1):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    dev_ms = timed_steps(step_device, wl.flush_l2, args.steps, barrier)
    # ---- end-to-end through the public API
    for _ in range(2):
        step_e2e()

    def e2e_stream_ms(steps):  # the same steps through the throughput form of the public call (upload of step i + 1 beside step i)
        barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for rec in wl.stream_e2e(steps):
            if dist is not None:
                gather_record(dist, rec, world, wl.n_pages, wl.n_tables)
        torch.cuda.current_stream().synchronize()
        b.record()
        barrier()
        return a.elapsed_time(b)

    e2e_call_ms = timed_loop(step_e2e, args.steps, barrier)  # one predict_pages call per step
    e2e_stream_ms(2)
    e2e_ms = e2e_stream_ms(args.steps)
    clocks = sampler.stop() if rank == 0 else None
    h2d, d2h = wl.e2e_bytes()
    api_boxes = sum(len(p["det"]) for p in wl.api_out)
    api_cells = sum(len(t[1]["polygons"]) for p in wl.api_out for t in p["tables"])
    # ---- per-kernel device times (CUDA events on the launching stream, same steps, separate pass).  The pass runs the step on ONE
    # stream: with the table branch on its own stream the events around a launch would also span the other branch's kernels
    two_streams, TWO_STREAMS = TWO_STREAMS, False
    for e in wl.engines:
        e.profile_begin()
    for _ in range(args.steps):
        wl.flush_l2()
        wl.step_device()
    recs = []
    for e in wl.engines:
        recs += e.profile_report()
    TWO_STREAMS = two_streams
    dev_ms, e2e_ms, e2e_call_ms = max_over_ranks(dist, [dev_ms, e2e_ms, e2e_call_ms], dev)

    peaks, peak_src = load_peaks()
    blocks = {}
    if not args.no_blocks:
        blocks["pp_rec"] = block_pp_rec(wl, rank, world, dist, args, barrier, peaks, peak_src)
        blocks["rec_sweep"] = block_rec_sweep(wl, rank, world, dist, args, barrier, peaks, peak_src)
        if FULL:
            blocks["lore"] = block_lore(wl, rank, world, dist, args, barrier, peaks, peak_src)
            blocks["lore_wireless"] = block_lore_wireless(wl, rank, local_rank, world, dist, args, barrier, peaks, peak_src)
            blocks["cascade_fp32x"] = block_fp32x(rank, local_rank, world, dist, args, barrier)

    if rank == 0:
        total_pages = wl.n_pages * world * args.steps
        agg = aggregate(recs)
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            cores = host_threads()
            arm = CpuArm(1 if FULL else 2)
            t0 = time.perf_counter()
            arm.step()
            secs = time.perf_counter() - t0
            cpu = {"value": arm.n_pages / secs, "unit": "pages/s", "cores": cores, "kind": "port", "sample": arm.sample() + f", one step, {secs:.1f} s"}
            for name, b in cpu_blocks().items():
                if name in blocks:
                    blocks[name]["cpu_baseline"] = {**b, "cores": cores}
        cfg = workload_config()
        cfg["stages"] = wl.stages + (["all_gather(packed results)"] if world > 1 else [])
        line = {
            "metric": "pages_per_sec", "value": total_pages / (dev_ms / 1e3), "unit": "pages/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16", "data": "synthetic", "config": cfg,
            "roofline": roofline_of(agg, peaks, peak_src, "full" if FULL else "ocr"), "cpu_baseline": cpu,
            "e2e": {"value": total_pages / (e2e_ms / 1e3), "unit": "pages/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms / args.steps,
                    "api": "pdf_table_b200.system.OcrSystemTask.predict_stream(iterator of numpy page batches [32,960,960,3] in pinned host "
                           "memory) -> per-page dicts (layout rows, boxes, strings, table cells) per batch; every step's upload and result "
                           "read-back is inside the timed loop, the upload of step i + 1 runs beside the device work of step i",
                    "per_call": {"value": total_pages / (e2e_call_ms / 1e3), "unit": "pages/s", "ms_per_step": e2e_call_ms / args.steps,
                                 "api": "one OcrSystemTask.predict_pages(batch) call per step (upload, device work and read-back in sequence)"},
                    "boxes_per_step": api_boxes, "cells_per_step": api_cells},
            "gpu_launches": int(launches_per_step * args.steps), "clocks": clocks, "kernels": kernel_table(agg, args.steps),
            "crops_per_page": CROPS_PER_PAGE, "crops_per_sec_in_cascade": wl.n_crops * world * args.steps / (dev_ms / 1e3),
            "collective": ("all_gather_into_tensor of the packed results inside every timed step (device and e2e legs)" if world > 1 else
                           "none at N=1 (single rank)"),
            "blocks": blocks,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
