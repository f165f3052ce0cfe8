"""world_size-2/3 gloo tests of the multi-GPU host logic (gather + de-interleave of cyclic bands)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from bhusie_b200.multi import BandLayout, gather_bands


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, h, w, band, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lay = BandLayout(h, band, world)
        full = torch.arange(h * w * 4, dtype=torch.float32).reshape(h, w, 4)       # what a 1-rank render would give
        local = torch.full((lay.max_local_rows, w, 4), -1.0)
        rows = torch.as_tensor(lay.rows_of(rank))
        local[: len(rows)] = full[rows]                                             # this rank's bands, band-major
        frame = torch.zeros((h, w, 4)) if rank == 0 else None
        out = gather_bands(local, lay, rank, frame)
        if rank == 0:
            q.put(bool(torch.equal(out, full)))
        else:
            assert out is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,h,w,band", [(2, 32, 5, 4), (2, 37, 3, 5), (3, 23, 4, 2)])
def test_gather_bands_gloo(world, h, w, band):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, h, w, band, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=5) is True


def test_gather_bands_single_rank():
    lay = BandLayout(6, 6, 1)
    local = torch.arange(6 * 2 * 4, dtype=torch.float32).reshape(6, 2, 4)
    assert gather_bands(local, lay, 0) is local
    frame = torch.zeros_like(local)
    assert torch.equal(gather_bands(local, lay, 0, frame), local)
