#!/usr/bin/env python
"""BASELINE configs[3]: the ConvNextViT document recogniser on 4096 text-line crops 32x320, crop-sharded over N B200s
(strong scaling: the 4096 crops are dealt round-robin, 4096 / N per GPU, no data-path collective; one all-gather of the
decoded ids at the end of a batch).  The metric is BASELINE.json's first one, text-line crops/sec:

    python tools/bench_rec.py [--crops 4096] [--steps 10] [--warmup 3]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/bench_rec.py

A step = rec pre-process (fused) + ConvNextViT + arg-max + CTC collapse of this rank's crops.  `value` has the uint8 crops
resident in HBM, `e2e` copies them from pinned host memory and reads the collapsed ids back inside the timed region.  Device
times are CUDA events, max over ranks; the L2 is flushed between timed steps.  Prints one JSON line on rank 0.  Not the
driver's bench (bench.py is); the CPU arm is `--impl reference` (oracle port on the host cores, a bounded sample)."""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pdf_table_b200 import sharding, synth, weights  # noqa: E402

CROP_H, CROP_W = 32, 320


def make_crops(n: int, first: int = 0, stride: int = 1) -> np.ndarray:
    """Crops first, first + stride, ... of the global list (64 distinct synthetic text lines, cycled)."""
    distinct = [synth.synthetic_text_crop(900000 + i, CROP_H, CROP_W) for i in range(64)]
    return np.stack([distinct[(first + k * stride) % 64] for k in range(n)])


def run_reference(args):
    """The reference's algorithm on the host cores: OCRRecognitionPreprocessor + ConvNextViT + post-processor (oracle port)."""
    from oracle import convnextvit_ref

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = {k: torch.from_numpy(v) for k, v in synth.convnext_vit_state_dict(0).items()}
    n = 32
    crops = list(make_crops(n))

    def step():
        for i in range(0, n, 16):
            convnextvit_ref.greedy_ids(convnextvit_ref.convnextvit_forward(sd, convnextvit_ref.preprocess(crops[i:i + 16])))

    step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    v = n / dt
    print(json.dumps({"impl": "reference", "metric": "text_line_crops_per_sec", "value": v, "unit": "crops/s", "n_gpus": args.gpus, "steps": args.steps,
                      "ms_per_step": dt * 1e3, "higher_is_better": True, "dtype": "f32", "data": "synthetic",
                      "config": {"workload": f"BASELINE configs[3]: ConvNextViT recogniser, {args.crops} crops {CROP_H}x{CROP_W}"},
                      "cpu_baseline": {"value": v, "unit": "crops/s", "cores": cores, "kind": "port",
                                       "sample": f"{n} of {args.crops} crops per step, oracle/ restatement in torch fp32"}}), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--crops", type=int, default=4096)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        if rank == 0:
            run_reference(args)
        return
    if not torch.cuda.is_available():
        raise SystemExit("bench_rec.py: no CUDA device -- the B200 arm has no CPU fallback (use --impl reference for the CPU arm)")
    from pdf_table_b200.engine import Engine

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

    counts = [len(range(r, args.crops, world)) for r in range(world)]  # round-robin deal of the crop list
    n = counts[rank]
    rec = Engine("convnext_vit", weights.pack_convnext_vit(synth.convnext_vit_state_dict(0)), device=local_rank)
    post = Engine("post", device=local_rank)
    crops_host = torch.from_numpy(make_crops(n, rank, world)).pin_memory()
    crops_dev = crops_host.to(dev)
    crops_stage = torch.empty_like(crops_dev)
    ids = torch.empty((n, 201), dtype=torch.int32, device=dev)
    ids_host = torch.empty((n, 201), dtype=torch.int32).pin_memory()
    len_host = torch.empty((n,), dtype=torch.int32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def step_device():
        rec.convnextvit_forward_u8(crops_dev, ids=ids)
        return post.ctc_collapse(ids)

    def step_e2e():
        crops_stage.copy_(crops_host, non_blocking=True)
        rec.convnextvit_forward_u8(crops_stage, ids=ids)
        out, ln, _ = post.ctc_collapse(ids)
        ids_host.copy_(out, non_blocking=True)
        len_host.copy_(ln, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    l0 = rec.launch_count + post.launch_count
    step_device()
    launches_per_step = rec.launch_count + post.launch_count - l0
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in evs:
        flush.fill_(1)
        a.record()
        step_device()
        b.record()
    barrier()
    dev_ms = sum(a.elapsed_time(b) for a, b in evs)
    for _ in range(2):
        step_e2e()
    barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(args.steps):
        step_e2e()
    b.record()
    barrier()
    e2e_ms = a.elapsed_time(b)
    rec.profile_begin()
    for _ in range(args.steps):
        flush.fill_(1)
        step_device()
    recs = rec.profile_report()
    t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out, ln, _ = step_device()
        none = {"boxes": torch.zeros((0, 1, 8), dtype=torch.float32, device=dev), "box_counts": torch.zeros((0,), dtype=torch.int32, device=dev)}
        gathered = sharding.all_gather_results({**none, "ids": out, "id_lens": ln}, [0] * world, counts)  # the one exchange of the path (rank-major order)
        assert gathered["id_lens"].numel() == args.crops
    dev_ms, e2e_ms = float(t[0]), float(t[1])
    if rank == 0:
        agg = {}
        for r in recs:
            k = agg.setdefault(r["kernel"], {"ms": 0.0, "flops": 0.0, "bytes": 0.0, "n": 0})
            k["ms"] += r["ms"]
            k["flops"] += r["flops"]
            k["bytes"] += r["bytes"]
            k["n"] += 1
        try:
            peaks = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")))
        except Exception:
            peaks = {}
        kernels = {k: {"ms_per_step": v["ms"] / args.steps, "launches_per_step": v["n"] / args.steps,
                       "tflops": (v["flops"] / (v["ms"] / 1e3) / 1e12) if v["flops"] else None, "gbs": v["bytes"] / (v["ms"] / 1e3) / 1e9}
                   for k, v in sorted(agg.items(), key=lambda kv: -kv[1]["ms"])}
        total = args.crops * args.steps
        line = {"metric": "text_line_crops_per_sec", "value": total / (dev_ms / 1e3), "unit": "crops/s", "n_gpus": world, "steps": args.steps,
                "warmup": max(args.warmup, 3), "ms_per_step": dev_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f16", "data": "synthetic",
                "config": {"workload": f"BASELINE configs[3]: ConvNextViT recogniser, {args.crops} synthetic text-line crops {CROP_H}x{CROP_W} "
                                       f"(3 chunks each), crop-sharded round-robin over {world} GPU(s)", "crops_per_gpu": counts,
                           "stages": ["rec_preprocess_u8(fused)", "convnextvit_forward+argmax", "ctc_collapse"],
                           "l2": "flushed between timed steps (256 MiB write)", "model_gflop_per_crop": rec.model_flops / max(n, 1) / 1e9},
                "e2e": {"value": total / (e2e_ms / 1e3), "unit": "crops/s", "h2d_bytes_per_step": int(crops_host.numel()) * world,
                        "d2h_bytes_per_step": int(ids_host.numel() * 4 + len_host.numel() * 4) * world, "ms_per_step": e2e_ms / args.steps},
                "gpu_launches": int(launches_per_step * args.steps), "kernels": kernels, "peaks": peaks}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
