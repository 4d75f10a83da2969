"""Synthetic scenes shaped like the matcher's outputs.

The RoMa network is out of scope, so its outputs are synthesised: pinhole cameras on orbit rings
looking at the origin, an analytic depth field per reference view back-projected and re-projected
into each neighbour to give the ``warp_AB`` channels in RoMa's normalised coordinates, plus noise
and gross outliers so every filter rejects a realistic fraction; certainty maps in two families:

* ``"T"`` tie-free: all values distinct f32 in [0.2, 0.9) -- bit-exact sampler parity is
  well-defined on these;
* ``"R"`` realistic: ``sigmoid`` of a smooth field + noise, floor-clamped at ``certainty_thresh``
  like reference core/pipeline.py:407, with a large saturated area above the 0.9 cap (ties).

Everything is generated with torch on the requested device (CPU for tests, GPU for the bench).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import numpy as np
import torch

from .core.camera_models import CameraRecord
from .core.config import ROMA_PRESETS


@dataclass
class SynthScene:
    cameras: List[CameraRecord]
    refs_local: List[int]          # indices into ``cameras``
    nn_table: np.ndarray           # [V, nn] int64 neighbour indices (self excluded)
    H: int                         # map resolution
    W: int
    h_match: int                   # match resolution (H_lr, W_lr)
    w_match: int
    nn: int

    @property
    def n_refs(self) -> int:
        return len(self.refs_local)

    @property
    def n_pairs(self) -> int:
        return self.n_refs * self.nn


def make_orbit_cameras(n_views: int, width: int = 1297, height: int = 840, radius: float = 4.0,
                       focal_frac: float = 0.74) -> List[CameraRecord]:
    """Pinhole cameras on stacked orbit rings, all looking at the origin.  Rings hold at most 200
    views so that nearest neighbours keep a ~2 degree baseline at any ``n_views``."""
    n_rings = max(1, math.ceil(n_views / 200))
    per_ring = math.ceil(n_views / n_rings)
    fx = focal_frac * width
    K = np.array([[fx, 0.0, width / 2.0], [0.0, fx, height / 2.0], [0.0, 0.0, 1.0]], dtype=np.float32)
    cams: List[CameraRecord] = []
    for v in range(n_views):
        ring, slot = divmod(v, per_ring)
        theta = 2.0 * math.pi * (slot + 0.37 * ring) / per_ring
        elev = 0.35 * (ring - 0.5 * (n_rings - 1)) + 0.08 * math.sin(3.0 * theta)
        centre = np.array([radius * math.cos(theta), elev, radius * math.sin(theta)], dtype=np.float64)
        fwd = -centre / np.linalg.norm(centre)
        right = np.cross(np.array([0.0, 1.0, 0.0]), fwd)
        right /= np.linalg.norm(right)
        down = np.cross(fwd, right)
        R = np.stack([right, down, fwd], axis=0)                 # world -> camera
        t = -R @ centre
        cams.append(CameraRecord.from_KRt(v, width, height, K, R, t, image_path=f"synthetic/{v:05d}.png"))
    return cams


def make_scene(n_views: int, setting: str = "fast", ref_fraction: float = 0.25, nn: int = 4,
               width: int = 1297, height: int = 840) -> SynthScene:
    h_lr, h_map = ROMA_PRESETS[setting]
    cams = make_orbit_cameras(n_views, width, height)
    centres = torch.from_numpy(np.stack([c.C for c in cams]).astype(np.float32))
    if n_views > 1:
        dist = torch.cdist(centres, centres)
        dist.fill_diagonal_(float("inf"))
        k = max(1, min(nn, n_views - 1))
        nn_table = torch.topk(dist, k, largest=False, dim=1).indices.numpy()
    else:
        nn_table = np.zeros((n_views, 0), dtype=np.int64)
    n_refs = max(1, min(n_views, int(round(n_views * ref_fraction))))
    refs = sorted(set(int(round(x)) for x in np.linspace(0, n_views - 1, n_refs)))
    return SynthScene(cameras=cams, refs_local=refs, nn_table=nn_table, H=h_map, W=h_map,
                      h_match=h_lr, w_match=h_lr, nn=nn_table.shape[1])


def _gen(device, seed: int) -> torch.Generator:
    g = torch.Generator(device=device)
    g.manual_seed(int(seed))
    return g


def synth_ref_inputs(scene: SynthScene, ref_pos: int, device="cpu", cert_family: str = "R",
                     seed: int = 0, noise_sigma: float = 1e-3, outlier_frac: float = 0.10,
                     certainty_thresh: float = 0.20) -> Dict[str, object]:
    """Matcher-shaped inputs for the ``ref_pos``-th reference view of ``scene``.

    Returns dict(cert f32 [nn,H,W], warp f32 [nn,H,W,4], image u8 [h_match,w_match,3],
    ref_index, nbr_indices).
    """
    dev = torch.device(device)
    H, W, hm, wm = scene.H, scene.W, scene.h_match, scene.w_match
    ref_i = scene.refs_local[ref_pos]
    nbrs = [int(j) for j in scene.nn_table[ref_i][: scene.nn]]
    camA = scene.cameras[ref_i]
    g = _gen(dev, seed * 1000003 + ref_i)

    # the matcher's reference grid (reference core/matcher.py:131-135)
    ys = torch.linspace(-1 + 1 / H, 1 - 1 / H, H, device=dev)
    xs = torch.linspace(-1 + 1 / W, 1 - 1 / W, W, device=dev)
    gy, gx = torch.meshgrid(ys, xs, indexing="ij")

    # pixel coords the path will decode (reference core/pipeline.py:655-656,681-683), in f64
    xA = (gx.double() + 1.0) * 0.5 * (wm - 1) * (camA.width / float(wm))
    yA = (gy.double() + 1.0) * 0.5 * (hm - 1) * (camA.height / float(hm))
    depth = 3.6 + 0.5 * torch.sin(2.1 * gx.double() + 0.3 * ref_i) * torch.cos(1.7 * gy.double())
    Kinv = torch.from_numpy(np.linalg.inv(camA.K.astype(np.float64))).to(dev)
    RA = torch.from_numpy(camA.R.astype(np.float64)).to(dev)
    tA = torch.from_numpy(camA.t.astype(np.float64)).to(dev).reshape(3)
    pix = torch.stack([xA, yA, torch.ones_like(xA)], dim=-1)                # [H,W,3]
    Xc = (pix @ Kinv.T) * depth.unsqueeze(-1)
    Xw = (Xc - tA) @ RA                                                       # R^T (Xc - t)

    warp = torch.empty((len(nbrs), H, W, 4), dtype=torch.float32, device=dev)
    cert = torch.empty((len(nbrs), H, W), dtype=torch.float32, device=dev)
    smooth_lo = 8
    for k, j in enumerate(nbrs):
        camB = scene.cameras[j]
        PB = torch.from_numpy(camB.P.astype(np.float64)).to(dev)
        q = Xw @ PB[:, :3].T + PB[:, 3]
        uB = q[..., 0] / q[..., 2]
        vB = q[..., 1] / q[..., 2]
        xBn = uB / (camB.width / float(wm)) / (0.5 * (wm - 1)) - 1.0
        yBn = vB / (camB.height / float(hm)) / (0.5 * (hm - 1)) - 1.0
        xBn = xBn + noise_sigma * torch.randn((H, W), generator=g, device=dev, dtype=torch.float64)
        yBn = yBn + noise_sigma * torch.randn((H, W), generator=g, device=dev, dtype=torch.float64)
        bad = torch.rand((H, W), generator=g, device=dev) < outlier_frac
        rx = torch.rand((H, W), generator=g, device=dev, dtype=torch.float64) * 2 - 1
        ry = torch.rand((H, W), generator=g, device=dev, dtype=torch.float64) * 2 - 1
        xBn = torch.where(bad, rx, xBn)
        yBn = torch.where(bad, ry, yBn)
        warp[k, ..., 0] = gx
        warp[k, ..., 1] = gy
        warp[k, ..., 2] = xBn.float()
        warp[k, ..., 3] = yBn.float()
        if cert_family == "R":
            lo = torch.randn((1, 1, smooth_lo, smooth_lo), generator=g, device=dev)
            field = torch.nn.functional.interpolate(lo, size=(H, W), mode="bicubic", align_corners=False)[0, 0]
            logit = 3.0 * field + 1.0 + 0.7 * torch.randn((H, W), generator=g, device=dev)
            c = torch.sigmoid(logit)
            c = torch.where(bad, c * 0.5, c)
            cert[k] = torch.clamp(c, min=certainty_thresh)
    if cert_family == "T":
        n = cert.numel()
        perm = torch.randperm(n, generator=g, device=dev)
        vals = 0.2 + 0.7 * (perm.double() + 0.5) / n
        cert = vals.float().reshape(cert.shape)
    elif cert_family != "R":
        raise ValueError("cert_family must be 'T' or 'R'")
    image = torch.randint(0, 256, (hm, wm, 3), generator=g, device=dev, dtype=torch.uint8)
    return {"cert": cert, "warp": warp, "image": image, "ref_index": ref_i, "nbr_indices": nbrs}
