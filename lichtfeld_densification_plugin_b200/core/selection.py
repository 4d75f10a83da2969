"""Camera selection on the device: the reference's ``core/selection.py`` signatures for the two pose-based functions.

* ``select_cameras_kcenters(flat_poses, k) -> List[int]``   reference core/selection.py:36-54
* ``nearest_neighbors(flat_poses, k) -> np.ndarray [n, k] int64``   reference core/selection.py:57-70
* ``_estimate_total_pairs``   reference core/pipeline.py:284-293 (host arithmetic on the neighbour table)

``select_cameras_by_visibility`` (core/selection.py:10-33) walks pycolmap's sparse point tracks and is out of scope.
No CPU fallback: both functions need the CUDA library and a device.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np
import torch

from .. import _native as N


def _device(device: Optional[torch.device]) -> torch.device:
    if not torch.cuda.is_available():
        raise N.NativeLibraryError("camera selection needs a CUDA device (there is no CPU fallback)")
    return torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")


def _poses_to_device(flat_poses, dev: torch.device) -> torch.Tensor:
    if isinstance(flat_poses, torch.Tensor):
        x = flat_poses.to(device=dev, dtype=torch.float32)
    else:
        x = torch.from_numpy(np.ascontiguousarray(np.asarray(flat_poses, dtype=np.float32))).to(dev)      # np.asarray(.., float32), :38
    if x.dim() != 2 or x.shape[1] != 16:
        raise ValueError("flat_poses must be [n, 16]")
    return x.contiguous()


def select_cameras_kcenters_device(flat_poses, k: int, device=None):
    """Device tensors (sorted centres [k] int32, pick order [k] int32); nothing synchronises."""
    dev = _device(device)
    x = _poses_to_device(flat_poses, dev)
    n = int(x.shape[0])
    k = max(1, min(int(k), n))                                       # :40
    lib = N.load()
    scratch = torch.empty_like(x)
    out_sorted = torch.empty((k,), dtype=torch.int32, device=dev)
    out_order = torch.empty((k,), dtype=torch.int32, device=dev)
    N.check(lib.ldp_select_kcenters(C.c_void_p(x.data_ptr()), n, k, C.c_void_p(scratch.data_ptr()), C.c_void_p(out_sorted.data_ptr()),
                                    C.c_void_p(out_order.data_ptr()), C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)),
            "ldp_select_kcenters")
    return out_sorted, out_order


def select_cameras_kcenters(flat_poses, k: int) -> List[int]:
    """k-centers selection using normalized camera poses (reference signature and return value)."""
    out_sorted, _ = select_cameras_kcenters_device(flat_poses, k)
    return [int(v) for v in out_sorted.cpu().tolist()]


def nearest_neighbors_device(flat_poses, k: int, device=None) -> torch.Tensor:
    dev = _device(device)
    x = _poses_to_device(flat_poses, dev)
    n = int(x.shape[0])
    if n <= 1:
        return torch.empty((n, 0), dtype=torch.int64, device=dev)    # :63-64
    k = max(1, min(int(k), n - 1))                                   # :65
    lib = N.load()
    out = torch.empty((n, k), dtype=torch.int64, device=dev)
    N.check(lib.ldp_nearest_neighbors(C.c_void_p(x.data_ptr()), n, k, C.c_void_p(out.data_ptr()),
                                      C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "ldp_nearest_neighbors")
    return out


def nearest_neighbors(flat_poses, k: int) -> np.ndarray:
    """Euclidean nearest neighbours across camera poses (reference signature and return value)."""
    return nearest_neighbors_device(flat_poses, k).cpu().numpy()


def _estimate_total_pairs(refs_local: List[int], nn_table: np.ndarray, img_ids: List[int], nns_per_ref: int) -> int:
    """reference core/pipeline.py:284-293"""
    return sum(sum(1 for n in nn_table[ref_idx][:nns_per_ref] if img_ids[n] != img_ids[ref_idx]) for ref_idx in refs_local)
