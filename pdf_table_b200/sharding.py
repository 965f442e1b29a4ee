"""Multi-GPU plumbing of the hot path: pages / crops are independent, so every rank (one process per
GPU) runs the whole cascade on its own contiguous shard with replicated weights, and the only exchange is ONE
all-gather per batch of the packed decoded results (boxes + counts + token ids + lengths).  No other collective.

The reference has no distributed inference at all (single process, cuda:0, batch 1: base_infer_task.py:69,
ocr_system_task.py:309-312); this module is the B200-side replacement for its serial page loop (cli/main.py:116-144).
"""
from __future__ import annotations

from typing import Dict, List, Sequence, Tuple

import torch


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block sharding; the first (n_items % world) ranks take one extra item."""
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_sizes(n_items: int, world: int) -> List[int]:
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


FIELDS = ("boxes", "box_counts", "ids", "id_lens")


def pack_results(res: Dict[str, torch.Tensor], max_pages: int, max_crops: int) -> torch.Tensor:
    """Fixed-size int32 record of one rank's results (float32 boxes are bit-cast), padded to the largest shard so a
    single all_gather_into_tensor suffices.  Layout: [n_pages, n_crops, boxes..., box_counts..., ids..., id_lens...]."""
    boxes, counts, ids, lens = (res[k] for k in FIELDS)
    n_pages, max_boxes = boxes.shape[0], boxes.shape[1]
    n_crops, t = ids.shape
    dev = boxes.device
    out = torch.zeros(2 + max_pages * max_boxes * 8 + max_pages + max_crops * t + max_crops, dtype=torch.int32, device=dev)
    out[0], out[1] = n_pages, n_crops
    o = 2
    out[o:o + n_pages * max_boxes * 8] = boxes.reshape(-1).view(torch.int32)
    o += max_pages * max_boxes * 8
    out[o:o + n_pages] = counts
    o += max_pages
    out[o:o + n_crops * t] = ids.reshape(-1)
    o += max_crops * t
    out[o:o + n_crops] = lens
    return out


def unpack_results(buf: torch.Tensor, max_pages: int, max_crops: int, max_boxes: int, t: int) -> Dict[str, torch.Tensor]:
    n_pages, n_crops = int(buf[0]), int(buf[1])
    o = 2
    boxes = buf[o:o + n_pages * max_boxes * 8].view(torch.float32).reshape(n_pages, max_boxes, 8)
    o += max_pages * max_boxes * 8
    counts = buf[o:o + n_pages]
    o += max_pages
    ids = buf[o:o + n_crops * t].reshape(n_crops, t)
    o += max_crops * t
    lens = buf[o:o + n_crops]
    return {"boxes": boxes, "box_counts": counts, "ids": ids, "id_lens": lens}


def all_gather_results(res: Dict[str, torch.Tensor], page_sizes: Sequence[int], crop_sizes: Sequence[int], group=None) -> Dict[str, torch.Tensor]:
    """ONE collective: every rank ends up with the results of all pages / crops in global order."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    max_pages, max_crops = max(page_sizes), max(crop_sizes)
    max_boxes, t = res["boxes"].shape[1], res["ids"].shape[1]
    mine = pack_results(res, max_pages, max_crops)
    gathered = torch.empty(world * mine.numel(), dtype=torch.int32, device=mine.device)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    parts = [unpack_results(gathered[r * mine.numel():(r + 1) * mine.numel()], max_pages, max_crops, max_boxes, t) for r in range(world)]
    return {k: torch.cat([p[k] for p in parts], 0) for k in FIELDS}
