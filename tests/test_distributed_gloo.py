"""CPU, world_size 2 over gloo: reference views shard contiguously and the final gather preserves order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from lichtfeld_densification_plugin_b200 import distributed as D


def test_shard_bounds_cover_and_order():
    for n in (0, 1, 5, 46, 250, 1000):
        for world in (1, 2, 3, 4, 8):
            got = []
            for r in range(world):
                lo, hi = D.shard_bounds(n, r, world)
                assert 0 <= lo <= hi <= n
                got += list(range(lo, hi))
            assert got == list(range(n))
            sizes = [D.shard_bounds(n, r, world)[1] - D.shard_bounds(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank: int, world: int, port: int, out_dir: str) -> None:
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        refs = list(range(11))
        mine = D.shard_refs(refs)
        # each "reference view" r yields r+1 points whose values encode (r, j)
        rows = [[r * 100 + j for _ in range(7)] for r in mine for j in range(r + 1)]
        pts = torch.tensor(rows, dtype=torch.float32).reshape(-1, 7)
        cap = 80                                   # padded like the device buffers
        xyz = torch.zeros((cap, 3)); rgb = torch.zeros((cap, 3)); err = torch.zeros((cap,))
        n = pts.shape[0]
        xyz[:n], rgb[:n], err[:n] = pts[:, 0:3], pts[:, 3:6], pts[:, 6]
        gx, gr, ge, counts = D.all_gather_points(xyz, rgb, err, n_valid=n)
        np.savez(os.path.join(out_dir, f"r{rank}.npz"), xyz=gx.numpy(), rgb=gr.numpy(), err=ge.numpy(), counts=counts.numpy(),
                 mine=np.asarray(mine))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(120)
def test_all_gather_points_world2(tmp_path):
    world = 2
    port = _free_port()
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    want = np.array([r * 100 + j for r in range(11) for j in range(r + 1)], dtype=np.float32)
    shards = []
    for rank in range(world):
        z = np.load(tmp_path / f"r{rank}.npz")
        shards.append(z["mine"].tolist())
        assert np.array_equal(z["err"], want)                   # single-process order, on every rank
        assert np.array_equal(z["xyz"][:, 0], want) and np.array_equal(z["rgb"][:, 2], want)
        assert z["counts"].sum() == want.size
    assert shards[0] + shards[1] == list(range(11))
