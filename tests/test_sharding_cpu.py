"""N>1 host logic on CPU: world_size-2 gloo processes shard a page/crop list, pack their (fake) decoded results and
exchange them with the single all-gather of the path."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pdf_table_b200 import sharding


def test_shard_ranges_cover_everything():
    for n in (0, 1, 7, 32, 256, 4096):
        for world in (1, 2, 3, 8):
            spans = [sharding.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sharding.shard_sizes(n, world)


def _fake_results(p0, p1, c0, c1):
    """Deterministic per-item results so the gathered tensors can be checked against a single-process run."""
    pages = torch.arange(p0, p1)
    boxes = (pages[:, None, None] * 1000 + torch.arange(5)[None, :, None] * 10 + torch.arange(8)[None, None, :]).float() + 0.5
    counts = (pages % 6).int()
    crops = torch.arange(c0, c1)
    ids = (crops[:, None] * 7 + torch.arange(11)[None, :]).int() % 97
    lens = (crops % 12).int()
    return {"boxes": boxes, "box_counts": counts, "ids": ids, "id_lens": lens}


def _fake_tables(t0, t1):
    tables = torch.arange(t0, t1)
    cells = (tables[:, None, None] * 100 + torch.arange(6)[None, :, None] + torch.arange(8)[None, None, :] * 0.25).float()
    logi = (tables[:, None, None] + torch.arange(6)[None, :, None] * torch.arange(4)[None, None, :]).float()
    return {"cells": cells, "cell_counts": (tables % 7).int(), "cell_logi": logi}


def _worker(rank, world, port, n_pages, n_crops):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        p0, p1 = sharding.shard_range(n_pages, rank, world)
        c0, c1 = sharding.shard_range(n_crops, rank, world)
        out = sharding.all_gather_results(_fake_results(p0, p1, c0, c1), sharding.shard_sizes(n_pages, world),
                                          sharding.shard_sizes(n_crops, world))
        want = _fake_results(0, n_pages, 0, n_crops)
        for k in sharding.FIELDS:
            assert torch.equal(out[k], want[k]), k
        assert set(out) == set(sharding.FIELDS)
        # the full record of SURVEY.md 8(e): boxes + token ids + table cells, ragged shards (5 tables over 2 ranks), same collective
        n_tables = 5
        t0, t1 = sharding.shard_range(n_tables, rank, world)
        out = sharding.all_gather_results({**_fake_results(p0, p1, c0, c1), **_fake_tables(t0, t1)}, sharding.shard_sizes(n_pages, world),
                                          sharding.shard_sizes(n_crops, world), table_sizes=sharding.shard_sizes(n_tables, world))
        want = {**want, **_fake_tables(0, n_tables)}
        for k in sharding.FIELDS + sharding.TABLE_FIELDS:
            assert out[k].dtype == want[k].dtype and torch.equal(out[k], want[k]), k
    finally:
        dist.destroy_process_group()


def test_all_gather_results_world2_gloo():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, 7, 25), nprocs=2, join=True)


def test_pack_results_rejects_inconsistent_records():
    r = _fake_results(0, 3, 0, 4)
    with pytest.raises(ValueError):
        sharding.pack_results({**r, "cells": torch.zeros((1, 6, 8))}, 3, 4, 1)  # a partial set of table fields
    with pytest.raises(ValueError):
        sharding.pack_results(r, 2, 4)  # more pages than the padded record holds
    with pytest.raises(TypeError):
        sharding.pack_results({**r, "ids": r["ids"].long()}, 3, 4)
    buf = sharding.pack_results(r, 5, 6)
    back = sharding.unpack_results(buf, r, 5, 6)
    assert all(torch.equal(back[k], r[k]) for k in sharding.FIELDS)
