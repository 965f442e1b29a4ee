"""Multi-GPU plumbing of the hot path: pages / crops are independent, so every rank (one process per
GPU) runs the whole cascade on its own contiguous shard with replicated weights, and the only exchange is ONE
all-gather per batch of the packed decoded results (boxes + counts + token ids + lengths, and the table cells --
polygons + counts + logical coordinates -- when table structure ran).  No other collective.

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
TABLE_FIELDS = ("cells", "cell_counts", "cell_logi")  # table structure: polygons [tables, max_cells, 8] f32, counts, logi [.., 4] f32
_KIND = {"boxes": 0, "box_counts": 0, "ids": 1, "id_lens": 1, "cells": 2, "cell_counts": 2, "cell_logi": 2}  # rows = pages / crops / tables


def _fields(res: Dict[str, torch.Tensor]) -> Tuple[str, ...]:
    """The record's fields: the four detection / recognition ones always, the three table-structure ones when given."""
    has = [k in res for k in TABLE_FIELDS]
    if any(has) and not all(has):
        raise ValueError(f"table-structure results need all of {TABLE_FIELDS}")
    return FIELDS + (TABLE_FIELDS if all(has) else ())


def _row_len(t: torch.Tensor) -> int:
    n = 1
    for d in t.shape[1:]:
        n *= int(d)
    return n


def pack_results(res: Dict[str, torch.Tensor], max_pages: int, max_crops: int, max_tables: int = 0) -> torch.Tensor:
    """Fixed-size int32 record of one rank's results (float32 tensors are bit-cast), every field padded to the largest shard
    so a single all_gather_into_tensor suffices.  Layout: [n_pages, n_crops, n_tables, field blocks in _fields() order]; a
    field's block is max_rows x (product of its trailing dims) words, the first n_rows rows valid."""
    names = _fields(res)
    maxes = (max_pages, max_crops, max_tables)
    counts = [res["boxes"].shape[0], res["ids"].shape[0], res["cells"].shape[0] if "cells" in res else 0]
    for k in names:
        if res[k].shape[0] != counts[_KIND[k]] or counts[_KIND[k]] > maxes[_KIND[k]]:
            raise ValueError(f"field '{k}': {res[k].shape[0]} rows, expected {counts[_KIND[k]]} <= {maxes[_KIND[k]]}")
    dev = res["boxes"].device
    out = torch.zeros(3 + sum(maxes[_KIND[k]] * _row_len(res[k]) for k in names), dtype=torch.int32, device=dev)
    out[0], out[1], out[2] = counts
    o = 3
    for k in names:
        t = res[k].contiguous()
        if t.dtype == torch.float32:
            t = t.view(torch.int32)
        elif t.dtype != torch.int32:
            raise TypeError(f"field '{k}' must be float32 or int32, got {t.dtype}")
        out[o:o + t.numel()] = t.reshape(-1)
        o += maxes[_KIND[k]] * _row_len(res[k])
    return out


def unpack_results(buf: torch.Tensor, like: Dict[str, torch.Tensor], max_pages: int, max_crops: int, max_tables: int = 0,
                   counts: Sequence[int] = ()) -> Dict[str, torch.Tensor]:
    """Inverse of pack_results for one rank's record; `like` = this rank's own results (trailing dims and dtypes are the
    same on every rank).  `counts` = the record's three header words when the caller has already read them."""
    names = _fields(like)
    maxes = (max_pages, max_crops, max_tables)
    counts = [int(c) for c in counts] if len(counts) == 3 else [int(buf[0]), int(buf[1]), int(buf[2])]
    out = {}
    o = 3
    for k in names:
        row, n = _row_len(like[k]), counts[_KIND[k]]
        t = buf[o:o + n * row]
        if like[k].dtype == torch.float32:
            t = t.view(torch.float32)
        out[k] = t.reshape((n,) + tuple(like[k].shape[1:]))
        o += maxes[_KIND[k]] * row
    return out


def all_gather_results(res: Dict[str, torch.Tensor], page_sizes: Sequence[int], crop_sizes: Sequence[int], group=None,
                       table_sizes: Sequence[int] = ()) -> Dict[str, torch.Tensor]:
    """ONE collective: every rank ends up with the results of all pages / crops (/ tables) in global order (rank-major)."""
    import torch.distributed as dist

    world = dist.get_world_size(group)
    max_pages, max_crops, max_tables = max(page_sizes), max(crop_sizes), max(table_sizes) if len(table_sizes) else 0
    mine = pack_results(res, max_pages, max_crops, max_tables)
    gathered = torch.empty(world * mine.numel(), dtype=torch.int32, device=mine.device)
    dist.all_gather_into_tensor(gathered, mine, group=group)
    heads = gathered.view(world, mine.numel())[:, :3].cpu().tolist()  # ONE device -> host read for all ranks' row counts
    parts = [unpack_results(gathered[r * mine.numel():(r + 1) * mine.numel()], res, max_pages, max_crops, max_tables, heads[r]) for r in range(world)]
    return {k: torch.cat([p[k] for p in parts], 0) for k in _fields(res)}
