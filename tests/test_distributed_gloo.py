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


def _worker_sharded_write(rank, world, port, tmp, q):
    import numpy as np
    import torch
    import torch.distributed as dist
    from lichtfeld_densification_plugin_b200 import distributed as D
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        rs = np.random.RandomState(5)
        xyz_all = rs.standard_normal((1000, 3)).astype(np.float32)
        rgb_all = rs.randint(0, 256, size=(1000, 3)).astype(np.uint8)
        bounds = [0, 377, 1000]                                   # uneven shards, as kept-point counts are
        lo, hi = bounds[rank], bounds[rank + 1]
        counts, offs = D.exchange_counts(torch.tensor([hi - lo]))
        assert counts.tolist() == [377, 623] and offs.tolist() == [0, 377]
        path = os.path.join(tmp, "sharded.ply")
        D.write_ply_sharded(path, xyz_all[lo:hi], rgb_all[lo:hi], int(offs[rank]), int(counts.sum()), rank)
        # the same files from pre-packed records (what the device packers hand over), and the points3D.bin variant
        from lichtfeld_densification_plugin_b200.core import writers as W
        rec = np.empty(hi - lo, dtype=W._PLY_VERTEX)
        rec["x"], rec["y"], rec["z"] = xyz_all[lo:hi, 0], xyz_all[lo:hi, 1], xyz_all[lo:hi, 2]
        rec["r"], rec["g"], rec["b"] = rgb_all[lo:hi, 0], rgb_all[lo:hi, 1], rgb_all[lo:hi, 2]
        path2 = os.path.join(tmp, "sharded_records.ply")
        D.write_ply_sharded(path2, None, None, int(offs[rank]), int(counts.sum()), rank, records=rec.view(np.uint8))
        brec = np.empty(hi - lo, dtype=W._BIN_POINT)
        brec["id"] = np.arange(lo + 1, hi + 1, dtype=np.uint64)
        brec["x"], brec["y"], brec["z"] = xyz_all[lo:hi, 0], xyz_all[lo:hi, 1], xyz_all[lo:hi, 2]
        brec["r"], brec["g"], brec["b"] = rgb_all[lo:hi, 0], rgb_all[lo:hi, 1], rgb_all[lo:hi, 2]
        brec["err"] = 0.0
        path3 = os.path.join(tmp, "sharded.bin")
        D.write_points3D_bin_sharded(path3, brec.view(np.uint8), int(offs[rank]), int(counts.sum()), rank)
        if rank == 0:
            ref = os.path.join(tmp, "whole.ply")
            W.write_ply(ref, xyz_all, rgb_all)
            refb = os.path.join(tmp, "whole.bin")
            W.write_points3D_bin(refb, xyz_all, rgb_all)
            want = open(ref, "rb").read()
            q.put(open(path, "rb").read() == want and open(path2, "rb").read() == want
                  and open(path3, "rb").read() == open(refb, "rb").read())
    finally:
        dist.destroy_process_group()


def test_counts_exchange_and_sharded_ply_world2(tmp_path):
    """World size 2, gloo: the ranks exchange only their kept-point counts and each writes its slice of one PLY file,
    which is byte-identical to the single-process writer's file for the rank-order concatenation."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sharded_write, args=(r, 2, port, str(tmp_path), q)) for r in range(2)]
    for p_ in procs:
        p_.start()
    for p_ in procs:
        p_.join(120)
        assert p_.exitcode == 0
    assert q.get(timeout=10) is True
