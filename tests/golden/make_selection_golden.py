"""Golden vectors for pair generation from the LIVE, UNMODIFIED reference core/selection.py.

    python tests/golden/make_selection_golden.py        (build container only: needs /root/reference)

Camera sets are regenerated from seeds by the tests (``selection_cases``); stored are the reference's outputs: the sorted
k-centres, the nearest-neighbour table, and - because torch.cdist's distances carry ~1e-3 of sgemm cancellation error,
which decides the order of near-equidistant neighbours - the reference's own k + 1 smallest distances per view, so that
the tests can tell a tie from a mismatch.
"""
from __future__ import annotations

import importlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
from lichtfeld_densification_plugin_b200 import synth  # noqa: E402


def selection_cases():
    """name -> (flat poses [n,16] float64 as CameraRecord.flat_pose() gives them, k for k-centres, k for neighbours)"""
    out = {}
    # small symmetric rings: rows of exactly equal norm (the first k-centres pick is np.einsum's rounding) and fewer than 26 views
    # (torch.cdist's direct kernel instead of the matrix product)
    for name, nv, kc, kn in (("ring1000", 1000, 250, 4), ("ring185", 185, 46, 4), ("ring40", 40, 32, 8),
                             ("ring24", 24, 3, 4), ("ring12", 12, 6, 4), ("ring25", 25, 25, 8), ("ring26", 26, 1, 8)):
        scene = synth.make_scene(nv, "turbo", 0.25, 4)
        out[name] = (np.stack([c.flat_pose() for c in scene.cameras], 0), kc, kn)
    rs = np.random.RandomState(77)
    for name, nv, kc, kn in (("random300", 300, 100, 10), ("random2", 2, 5, 3), ("random1", 1, 1, 3), ("random1500", 1500, 40, 16)):
        T = np.tile(np.eye(4).reshape(1, 16), (nv, 1))
        for i in range(nv):
            q, _ = np.linalg.qr(rs.standard_normal((3, 3)))
            M = np.eye(4)
            M[:3, :3] = q.astype(np.float32)
            M[:3, 3] = (rs.standard_normal(3) * 5).astype(np.float32)
            T[i] = M.reshape(-1)
        out[name] = (T, kc, kn)
    return out


def main() -> None:
    ref_import.import_reference(full_pipeline=False)
    sel = importlib.import_module("core.selection")
    torch.set_num_threads(1)
    out = {}
    for name, (flat, kc, kn) in selection_cases().items():
        out[f"{name}_centers"] = np.asarray(sel.select_cameras_kcenters(flat, kc), dtype=np.int64)
        nn = sel.nearest_neighbors(flat, kn)
        out[f"{name}_nn"] = nn
        n = flat.shape[0]
        if n > 1:
            m = torch.from_numpy(flat.astype(np.float32))
            D = torch.cdist(m, m, p=2)
            D.fill_diagonal_(float("inf"))
            kk = min(nn.shape[1] + 1, n)
            out[f"{name}_dist"] = torch.topk(D, kk, largest=False, dim=1)[0].numpy()
    out["versions"] = np.array(f"numpy {np.__version__} torch {torch.__version__}")
    path = os.path.join(HERE, "selection.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
