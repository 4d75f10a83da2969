"""Multi-GPU plumbing: reference views shard across ranks; the ranks exchange point COUNTS, not points.

The path partitions naturally -- reference views are independent units once the sampler's RNG is keyed per
view (Philox stream = global reference index) -- so each rank processes a contiguous slice of ``refs_local``.
Concatenating the ranks' outputs in rank order reproduces the single-GPU output order exactly, so the only thing a rank
needs from the others is how many points they kept: ``exchange_counts`` (8 bytes per rank) gives every rank its global
row offset, and ``write_ply_sharded`` lets each rank write its own slice of the output file (the reference's single
process concatenates and writes everything itself, core/pipeline.py:914-928, core/writers.py:29-46).  Shipping every
point to every rank (``all_gather_points``, 28 B/point) is available for callers that want the whole cloud on each GPU,
but it grows with the number of ranks and is not needed to produce the file.

Works with any ``torch.distributed`` backend (NCCL over NVLink on the GPU box, gloo in the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_refs: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of the reference list for ``rank``: floor(rank*n/world) .. floor((rank+1)*n/world)."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    return (rank * n_refs) // world, ((rank + 1) * n_refs) // world


def shard_refs(refs: Sequence[int], rank: Optional[int] = None, world: Optional[int] = None) -> List[int]:
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    lo, hi = shard_bounds(len(refs), rank, world)
    return list(refs[lo:hi])


def exchange_counts(n_points: torch.Tensor, group=None) -> Tuple[torch.Tensor, torch.Tensor]:
    """All-gather of one int64 per rank (the rank's kept-point count, a 1-element tensor on the compute device; no host
    synchronisation).  Returns (counts [world], exclusive prefix = global row offset of each rank's first point)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        counts = n_points.reshape(1).to(torch.int64)
        return counts, torch.zeros_like(counts)
    world = dist.get_world_size(group)
    counts = torch.empty(world, dtype=torch.int64, device=n_points.device)
    dist.all_gather_into_tensor(counts, n_points.reshape(1).to(torch.int64), group=group)
    return counts, torch.cumsum(counts, 0) - counts


def _write_sharded(path_out: str, header: bytes, payload, byte_offset: int, total_bytes: int, rank: int, group) -> None:
    import os
    if rank == 0:
        with open(path_out, "wb") as f:
            f.write(header)
            f.truncate(len(header) + int(total_bytes))
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.barrier(group=group)                    # the file exists with its final size before anybody seeks into it
    fd = os.open(path_out, os.O_WRONLY)
    try:
        view = memoryview(payload).cast("B")
        pos, base = 0, len(header) + int(byte_offset)
        while pos < len(view):                       # a single write is capped at 0x7ffff000 bytes and may come back short
            n = os.pwrite(fd, view[pos:pos + (1 << 30)], base + pos)
            if n <= 0:
                raise OSError(f"short write to {path_out} at byte {base + pos}")
            pos += n
    finally:
        os.close(fd)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.barrier(group=group)


def write_ply_sharded(path_out: str, xyz, rgb_uint8, row_offset: int, total_rows: int, rank: int, group=None,
                      records=None) -> None:
    """Every rank writes its own rows of ONE binary PLY file, byte-identical to ``core.writers.write_ply`` of the
    rank-order concatenation: rank 0 writes the header for ``total_rows`` vertices, each rank writes 15-byte records at
    header + 15 * row_offset.  ``xyz`` / ``rgb_uint8``: this rank's numpy arrays - or ``records``: the rank's records
    already packed on the device (``output.ply_records(...).cpu().numpy()``, uint8 [15 * n]).  Needs a file system all
    ranks share."""
    import numpy as np

    from .core.writers import _PLY_VERTEX, ply_header

    if records is None:
        n = int(xyz.shape[0])
        rec = np.empty(n, dtype=_PLY_VERTEX)
        rec["x"], rec["y"], rec["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
        rec["r"], rec["g"], rec["b"] = rgb_uint8[:, 0], rgb_uint8[:, 1], rgb_uint8[:, 2]
        payload = rec.view(np.uint8).reshape(-1)
    else:
        payload = np.ascontiguousarray(records, dtype=np.uint8).reshape(-1)
    _write_sharded(path_out, ply_header(int(total_rows)), payload, 15 * int(row_offset), 15 * int(total_rows), rank, group)


def write_points3D_bin_sharded(path_out: str, records, row_offset: int, total_rows: int, rank: int, group=None) -> None:
    """Same for COLMAP's points3D.bin (reference core/writers.py:15-26): ``records`` are this rank's 43-byte records
    packed with ``output.points3d_records(..., first_id=row_offset + 1)`` (ids are global row numbers + 1)."""
    import numpy as np
    payload = np.ascontiguousarray(records, dtype=np.uint8).reshape(-1)
    _write_sharded(path_out, np.uint64(int(total_rows)).tobytes(), payload, 43 * int(row_offset), 43 * int(total_rows), rank, group)


def all_gather_cloud(cloud, group=None, out=None, scratch: Optional[torch.Tensor] = None):
    """The final point all-gather, replacing the reference's single-process concatenation
    (core/pipeline.py:914-928): every rank contributes its ``output.PackedCloud`` (header with the device-side point count
    | xyz | rgb | err, 28 B/point, padded to ``cloud.capacity`` - the SAME capacity on every rank) in ONE collective;
    ``ldp_concat_points`` then squeezes the padding out in rank order, reading the counts from the gathered headers.
    Nothing synchronises with the host.  Returns (``PackedCloud`` with every rank's points in rank order = the single-GPU
    order, int64 device tensor [world + 1] of the ranks' global row offsets)."""
    from .output import ConcatPlan, PackedCloud
    dev = cloud.packed.device
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        off = torch.zeros((2,), dtype=torch.int64, device=dev)
        off[1:] = cloud.count
        return cloud, off
    world = dist.get_world_size(group)
    nbytes = cloud.packed.numel()
    if scratch is None:
        scratch = torch.empty((world, nbytes), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(scratch.view(-1), cloud.packed, group=group)
    blocks = [PackedCloud(cloud.capacity, dev, storage=scratch[r]) for r in range(world)]
    plan = ConcatPlan([b.xyz for b in blocks], [b.rgb for b in blocks], [b.err for b in blocks], [b.count for b in blocks],
                      cloud.capacity)
    if out is None:
        out = PackedCloud(world * cloud.capacity, dev)
    plan.run(out)
    out._plan, out._blocks = plan, scratch
    return out, plan.seg_offsets


class PeerClouds:
    """The final point all-gather as ONE kernel of ours over NVLink peer memory, instead of a library collective followed by a
    compaction.  Every rank keeps its ``output.PackedCloud`` AND the rank-ordered result cloud in symmetric memory
    (``torch.distributed._symmetric_memory``: the same allocations mapped into every rank's address space of the node).
    ``gather`` (push, the default) runs ``ldp_scatter_points``: the rank reads every peer's point count from the peer's
    header and writes its own rows - exactly that many, nothing padded or staged - into EVERY rank's result at the row the
    lower ranks' counts give, with posted 16-byte stores over NVLink / NVSwitch.  ``mode="pull"`` is the mirror image
    (``ldp_concat_points`` with the peers' clouds as sources: remote loads, each a round trip - slower, kept for comparison).
    Two device-side barriers of the symmetric-memory handle (signal pads, stream-ordered) bracket the kernel: every rank's
    cloud and count are complete before anybody looks at them, and every push has landed before anybody reads its result
    (or rewrites its cloud).  The host learns no count.
    Replaces reference core/pipeline.py:914-928 (single-process np.concatenate)."""

    def __init__(self, capacity: int, device, group=None) -> None:
        import numpy as np
        import torch.distributed._symmetric_memory as symm_mem
        from .output import ConcatPlan, PackedCloud
        self.group = group if group is not None else dist.group.WORLD
        self.world = dist.get_world_size(self.group)
        self.rank = dist.get_rank(self.group)
        dev = torch.device(device)
        nbytes = PackedCloud.nbytes(capacity)
        self.storage = symm_mem.empty(nbytes, dtype=torch.uint8, device=dev)
        self.storage.zero_()
        self.handle = symm_mem.rendezvous(self.storage, self.group)
        self.cloud = PackedCloud(capacity, dev, storage=self.storage)              # this rank's cloud: launches concatenate into it
        self.capacity = self.cloud.capacity
        self.peers = [self.cloud if p == self.rank else
                      PackedCloud(capacity, dev, storage=self.handle.get_buffer(p, (nbytes,), torch.uint8, 0))
                      for p in range(self.world)]
        self.plan = ConcatPlan([c.xyz for c in self.peers], [c.rgb for c in self.peers], [c.err for c in self.peers],
                               [c.count for c in self.peers], self.capacity)
        # the result, also symmetric: every rank writes its slice into every rank's copy
        out_bytes = PackedCloud.nbytes(self.world * self.capacity)
        self.out_storage = symm_mem.empty(out_bytes, dtype=torch.uint8, device=dev)
        self.out_storage.zero_()
        self.out_handle = symm_mem.rendezvous(self.out_storage, self.group)
        self.out = PackedCloud(self.world * self.capacity, dev, storage=self.out_storage)
        self.out_peers = [self.out if p == self.rank else
                          PackedCloud(self.world * self.capacity, dev, storage=self.out_handle.get_buffer(p, (out_bytes,), torch.uint8, 0))
                          for p in range(self.world)]
        tab = np.array([[c.count.data_ptr() for c in self.peers], [o.xyz.data_ptr() for o in self.out_peers],
                        [o.rgb.data_ptr() for o in self.out_peers], [o.err.data_ptr() for o in self.out_peers]], dtype=np.int64)
        self.tables = torch.from_numpy(tab).to(dev)
        self.seg_offsets = torch.zeros((self.world + 1,), dtype=torch.int64, device=dev)

    def gather(self, out=None, mode: str = "push"):
        """Every rank's points in rank order (= the single-GPU order) on THIS rank; returns (PackedCloud, int64 device tensor
        [world + 1] of the ranks' global row offsets).  Stream-ordered on the current stream; no host synchronisation.
        ``mode="push"`` fills the symmetric ``self.out`` (``out`` is ignored); ``mode="pull"`` fills ``out``."""
        import ctypes as C
        from . import _native as N
        from .output import PackedCloud
        if mode == "pull":
            if out is None:
                out = PackedCloud(self.world * self.capacity, self.cloud.packed.device)
            self.handle.barrier(channel=0)
            self.plan.run(out)
            self.handle.barrier(channel=1)
            return out, self.plan.seg_offsets
        lib = N.load()
        stream = torch.cuda.current_stream(self.cloud.packed.device).cuda_stream
        t = self.tables
        self.handle.barrier(channel=0)
        N.check(lib.ldp_scatter_points(C.c_void_p(self.cloud.xyz.data_ptr()), C.c_void_p(self.cloud.rgb.data_ptr()),
                                       C.c_void_p(self.cloud.err.data_ptr()), C.c_void_p(t[0].data_ptr()), self.rank, self.world,
                                       self.capacity, C.c_void_p(t[1].data_ptr()), C.c_void_p(t[2].data_ptr()),
                                       C.c_void_p(t[3].data_ptr()), self.out.capacity, C.c_void_p(self.seg_offsets.data_ptr()),
                                       C.c_void_p(self.out.count.data_ptr()), C.c_void_p(stream)), "ldp_scatter_points")
        self.handle.barrier(channel=1)
        return self.out, self.seg_offsets


def all_gather_points(xyz: torch.Tensor, rgb: torch.Tensor, err: torch.Tensor, n_valid: Optional[int] = None,
                      group=None) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """Order-preserving all-gather of per-rank packed points, exact-size results.

    ``xyz`` [cap,3], ``rgb`` [cap,3], ``err`` [cap] hold ``n_valid`` points (default: all rows).  Returns the
    concatenation in rank order plus the per-rank counts.  On CUDA tensors this is ``all_gather_cloud`` (one padded
    collective + the device-side concatenation) followed by the single host read that sizing the returned tensors needs;
    callers that can keep capacity-sized buffers use ``all_gather_cloud`` directly and never synchronise.  On CPU tensors
    (gloo, the host-logic tests) the same protocol runs with torch indexing.
    """
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        n = xyz.shape[0] if n_valid is None else int(n_valid)
        return xyz[:n], rgb[:n], err[:n], torch.tensor([n], dtype=torch.int64, device=xyz.device)
    world = dist.get_world_size(group)
    n = xyz.shape[0] if n_valid is None else int(n_valid)
    dev = xyz.device
    if xyz.is_cuda:
        from .output import PackedCloud
        caps = torch.tensor([int(xyz.shape[0])], dtype=torch.int64, device=dev)
        dist.all_reduce(caps, op=dist.ReduceOp.MAX, group=group)          # one capacity for every rank (set-up, not data path)
        cloud = PackedCloud(int(caps.item()), dev)
        cloud.xyz[:n], cloud.rgb[:n], cloud.err[:n] = xyz[:n], rgb[:n], err[:n]
        cloud.count.fill_(n)
        total, offsets = all_gather_cloud(cloud, group=group)
        k = int(offsets[-1].item())
        return total.xyz[:k], total.rgb[:k], total.err[:k], offsets[1:] - offsets[:-1]
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([n], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, mine, group=group)
    cmax = max(int(counts.max()), 1)
    fused = torch.zeros((cmax, 7), dtype=torch.float32, device=dev)
    fused[:n, 0:3] = xyz[:n]
    fused[:n, 3:6] = rgb[:n]
    fused[:n, 6] = err[:n]
    gathered = torch.empty((world * cmax * 7,), dtype=torch.float32, device=dev)
    dist.all_gather_into_tensor(gathered, fused.reshape(-1), group=group)
    gathered = gathered.view(world, cmax, 7)
    cl = counts.tolist()
    allp = torch.cat([gathered[r, :cl[r]] for r in range(world)], dim=0)
    return allp[:, 0:3].contiguous(), allp[:, 3:6].contiguous(), allp[:, 6].contiguous(), counts
