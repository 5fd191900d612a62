"""CPU, world_size 2, gloo: the spatially tiled select protocol (histogram all-reduce between
radix rounds + min all-reduce) yields on every rank the threshold of the WHOLE unit."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import pic_oracle as po
from _common import trained_like


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, units, n, prs, ret):
    import sys
    here = os.path.dirname(os.path.abspath(__file__))
    for p in (os.path.dirname(here), here, os.path.join(os.path.dirname(here), "oracle")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import pic_b200
        from pic_b200.distributed import allreduce_min_u32, shard_range, tiled_select_threshold
        from _numpy_tile_backend import NumpyTileBackend

        rng = np.random.default_rng(1234)
        std = trained_like(rng, (units, n))[3]
        std[:, ::3] = np.round(std[:, ::3] * 8) / 8          # ties across the tile boundary
        std[1, :] = 0.5                                       # all-equal unit
        # ragged row bands: rank 0 gets 1/3 of the columns, rank 1 the rest
        cut = n // 3
        tile = std[:, :cut] if rank == 0 else std[:, cut:]
        q = torch.tensor([pic_b200.ops.pr_to_q01(p) for p in prs], dtype=torch.float32)
        thr = tiled_select_threshold(None, units, n, q, backend=NumpyTileBackend(tile, units))
        _, ref = po.channel_mask(std, prs)
        ok = np.array_equal(thr.numpy(), ref, equal_nan=True)
        # unsigned MIN through the signed collective
        keys = torch.tensor([0x80000001 - (1 << 32) if rank == 0 else 5, 7 if rank == 0 else -2], dtype=torch.int32)
        mn = allreduce_min_u32(keys)
        ok = ok and mn.tolist() == [5, 7]
        b, e = shard_range(units, rank, world)
        ret[rank] = (bool(ok), thr.numpy().tolist(), ref.tolist(), (b, e))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_tiled_select_two_ranks_gloo():
    po.build()
    units, n = 6, 6000
    prs = [0.5, 5, 9.9999, 0, 10, 1e-4]
    mgr = mp.get_context("spawn").Manager()   # no fork() of a process that has run host threads
    ret = mgr.dict()
    port = _free_port()
    mp.spawn(_worker, args=(2, port, units, n, prs, ret), nprocs=2, join=True)
    assert set(ret.keys()) == {0, 1}
    for r in (0, 1):
        ok, thr, ref, span = ret[r]
        assert ok, (r, thr, ref)
    assert ret[0][1] == ret[1][1] or np.array_equal(np.asarray(ret[0][1]), np.asarray(ret[1][1]), equal_nan=True)
    assert ret[0][3] == (0, 3) and ret[1][3] == (3, 6)
