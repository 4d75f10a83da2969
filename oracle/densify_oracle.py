"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's post-matching densification path.

This file is the *oracle* for the CUDA path: a numpy / torch-CPU restatement of what the reference
plugin (shadygm/Lichtfeld-Densification-Plugin v0.8.3) computes between the RoMa matcher and the
point accumulator.  It is a checker, never a fallback: only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.  The product package
must not.

Parity status: the reference ships **no** tests, golden vectors or fixtures for this path
(SURVEY.md section 4), so the restatement is pinned against the *live* reference instead: it was
checked bit-for-bit against the unmodified ``/root/reference/core`` modules imported in the build
container (``tests/test_oracle_vs_reference.py``), and against golden vectors generated from that
live import (``tests/golden/make_golden.py`` -> ``tests/golden/*.npz``), which travel to the GPU box.

Every function cites the reference lines it restates (paths relative to the reference root).
The restatement deliberately keeps the reference's *operation structure* (legacy
``RandomState.choice``, a Python walk over ``argsort(-weights)``, LAPACK batched SVD, sgemm
projections) so that timing it is a fair "reference CPU path" baseline, and keeps its dtype flow
(f32 geometry, f64 Sampson, f64 bilinear weights) so results are bit-identical on the same box.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

__all__ = [
    "OracleCamera",
    "OracleConfig",
    "legacy_choice_no_replace",
    "philox4x32_10",
    "philox_uniforms",
    "sample_weights",
    "coverage_picks",
    "select_samples",
    "resize_mask_nearest",
    "warp_mask_nearest",
    "certainty_prologue",
    "certainty_prologue_torch",
    "best_neighbour",
    "decode_samples",
    "bilinear_colour",
    "fundamental_matrix",
    "sampson_distance",
    "dlt_points",
    "reprojection_error",
    "in_front",
    "parallax_ok",
    "triangulate_ref",
    "to_uint8_rgb",
    "ply_bytes",
    "points3d_bin_bytes",
]


# ----------------------------------------------------------------------------------------------
# plain-data stand-ins for the reference's types
# ----------------------------------------------------------------------------------------------
@dataclass
class OracleCamera:
    """Per-view constants, dtype/shape exactly as core/camera_models.py:10-21 holds them
    (K,R [3,3] f32; t [3,1] f32; P [3,4] f32 = K@[R|t]; C [3] f32 = -R^T t; densify.py:226-230)."""

    uid: int
    width: int
    height: int
    K: np.ndarray
    R: np.ndarray
    t: np.ndarray
    P: np.ndarray
    C: np.ndarray


@dataclass
class OracleConfig:
    """The scalars of core/config.py:7-26 that the hot path reads (+ the matcher constants of
    core/pipeline.py:99-105 / core/matcher.py:92-94)."""

    matches_per_ref: int = 10000
    reproj_thresh: float = 0.8
    sampson_thresh: float = 5.0
    min_parallax_deg: float = 0.5
    no_filter: bool = False
    sample_cap: float = 0.9      # matcher.sample_thresh, core/matcher.py:92
    border: int = 2              # core/pipeline.py:646
    tiles: int = 24              # core/pipeline.py:647
    w_match: int = 512
    h_match: int = 512


# ----------------------------------------------------------------------------------------------
# sampler  (core/sampling.py:8-53)
# ----------------------------------------------------------------------------------------------
def legacy_choice_no_replace(p32: np.ndarray, size: int, uniforms: np.ndarray) -> Tuple[np.ndarray, int, int]:
    """numpy's legacy ``RandomState.choice(n, size, replace=False, p=p)`` restated on an explicit
    uniform stream (numpy 2.3 ``numpy/random/mtrand.pyx``, pinned by the reference at
    ``pyproject.toml:12``; called from core/sampling.py:32).

    ``uniforms`` must be the ``random_sample`` stream of the generator ``choice`` would have used
    (``np.random.RandomState(seed).random_sample(k)``).  Returns ``(found, n_consumed, n_rounds)``;
    ``found`` is in numpy's draw order.  Raises the same ValueErrors numpy raises.
    """
    p = np.array(p32, dtype=np.float64)            # PyArray_FROM_OTF(p, NPY_DOUBLE)
    n = p.shape[0]
    if np.isnan(p).any():
        raise ValueError("probabilities contain NaN")
    if (p < 0).any():
        raise ValueError("probabilities are not non-negative")
    atol = np.sqrt(np.finfo(np.float64).eps)
    if isinstance(p32, np.ndarray) and np.issubdtype(p32.dtype, np.floating):
        atol = max(atol, np.sqrt(np.finfo(p32.dtype).eps))
    if abs(float(np.sum(p)) - 1.0) > atol:         # numpy uses a Kahan sum; same verdict away from atol
        raise ValueError("probabilities do not sum to 1")
    if size > n:
        raise ValueError("Cannot take a larger sample than population when 'replace=False'")
    if np.count_nonzero(p > 0) < size:
        raise ValueError("Fewer non-zero entries in p than size")

    found = np.zeros(size, dtype=np.int64)
    have = 0
    used = 0
    rounds = 0
    while have < size:
        want = size - have
        if used + want > uniforms.shape[0]:
            raise RuntimeError("explicit uniform stream exhausted")
        x = uniforms[used:used + want]
        used += want
        rounds += 1
        if have > 0:
            p[found[:have]] = 0.0
        cdf = np.cumsum(p)                         # sequential f64 running sum
        cdf /= cdf[-1]
        hit = cdf.searchsorted(x, side="right")    # first i with cdf[i] > x
        _, first_pos = np.unique(hit, return_index=True)
        first_pos.sort()
        hit = hit.take(first_pos)                  # first occurrence of each value, draw order
        found[have:have + hit.size] = hit
        have += hit.size
    return found, used, rounds


# ----------------------------------------------------------------------------------------------
# production uniforms: the product's counter-based stream restated on the host, so that a Philox-mode GPU
# run can be compared value for value with ``legacy_choice_no_replace`` fed the same doubles.
# (Not reference code: the reference draws from numpy's global MT19937, core/sampling.py:32.  What is
# restated here is lichtfeld_densification_plugin_b200/csrc/ldp_device.cuh:philox4x32_10 / philox_uniform.)
# ----------------------------------------------------------------------------------------------
_PHILOX_M0, _PHILOX_M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
_PHILOX_W0, _PHILOX_W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(counter: np.ndarray, key: Sequence[int]) -> np.ndarray:
    """Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3", SC'11;
    Random123 ``philox4x32_R(10, ctr, key)``).  ``counter``: uint32 [..., 4]; ``key``: two 32-bit words.
    Returns uint32 [..., 4].  Known-answer vectors of Random123's ``kat_vectors`` are in tests/test_oracle_philox.py."""
    c = np.asarray(counter, dtype=np.uint32).astype(np.uint64)
    c0, c1, c2, c3 = c[..., 0], c[..., 1], c[..., 2], c[..., 3]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    sh = np.uint64(32)
    for _ in range(10):
        p0 = _PHILOX_M0 * c0                     # 32 x 32 -> 64 bit products
        p1 = _PHILOX_M1 * c2
        hi0, lo0 = p0 >> sh, p0 & mask
        hi1, lo1 = p1 >> sh, p1 & mask
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + _PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + _PHILOX_W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


PHILOX_DOMAIN_TAG = 0x4C445031                   # "LDP1": fourth counter word of the sampler's stream


def philox_uniforms(seed: int, stream: int, n: int, first: int = 0) -> np.ndarray:
    """Draws ``first .. first + n - 1`` of view ``stream``'s production stream, as the kernels build them
    (ldp_device.cuh:philox_uniform): draw d uses Philox counter (d >> 1, 0, stream, "LDP1") under key
    (seed & 2^32-1, seed >> 32); even d takes output words (0, 1), odd d words (2, 3); the double is numpy's
    ``random_sample`` construction ((a >> 5) * 2^26 + (b >> 6)) / 2^53, so both modes share one lattice."""
    d = np.arange(first, first + n, dtype=np.uint64)
    ctr = np.zeros((n, 4), dtype=np.uint32)
    ctr[:, 0] = (d >> np.uint64(1)).astype(np.uint32)
    ctr[:, 2] = np.uint32(int(stream) & 0xFFFFFFFF)
    ctr[:, 3] = np.uint32(PHILOX_DOMAIN_TAG)
    r = philox4x32_10(ctr, (int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF))
    odd = (d & np.uint64(1)).astype(bool)
    a = np.where(odd, r[:, 2], r[:, 0]).astype(np.uint64)
    b = np.where(odd, r[:, 3], r[:, 1]).astype(np.uint64)
    return ((a >> np.uint64(5)).astype(np.float64) * 67108864.0 + (b >> np.uint64(6)).astype(np.float64)) / 9007199254740992.0


def sample_weights(best_cert: torch.Tensor, cap: float, border: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """Capped, border-masked f32 weights and their torch-CPU f32 sum (core/sampling.py:12-14,23-26).

    The value of ``s`` depends on torch's CPU reduction order (thread count / SIMD width)
    -- parity runs hand this exact ``s`` to the CUDA path as ``weight_sum_override``.
    """
    cert = torch.clamp(best_cert.clone().to("cpu"), max=cap)
    H, W = cert.shape
    rows = torch.arange(H).view(H, 1).expand(H, W)
    cols = torch.arange(W).view(1, W).expand(H, W)
    keep = (cols >= border) & (cols <= W - 1 - border) & (rows >= border) & (rows <= H - 1 - border)
    w = (cert * keep.float()).reshape(-1)
    return w, w.sum()


def coverage_picks(p: np.ndarray, H: int, W: int, tiles: int, budget: int) -> np.ndarray:
    """Per-tile best pixel, walking pixels in descending weight (core/sampling.py:34-50).

    ``np.argsort(-p)`` is an unstable sort: among equal weights the visiting order -- hence the
    pick inside a tile whose maximum is tied -- is implementation-defined.
    """
    tile = max(1, W // tiles)
    flat = np.arange(H * W)
    key = ((flat % W) // tile) * 100000 + ((flat // W) // tile)     # gx*100000 + gy
    picked: List[int] = []
    taken = set()
    for i in np.argsort(-p):
        if p[i] <= 0:
            break
        b = int(key[i])
        if b in taken:
            continue
        taken.add(b)
        picked.append(i)
        if len(picked) >= budget:
            break
    return np.asarray(picked, dtype=np.int64)


def select_samples(best_cert: torch.Tensor, M: int, cap: float = 0.9, border: int = 2, tiles: int = 24,
                   no_filter: bool = False, rng: Optional[np.random.RandomState] = None,
                   uniforms: Optional[np.ndarray] = None, taps: Optional[dict] = None,
                   s_override: Optional[np.float32] = None) -> np.ndarray:
    """core/sampling.py:8-53.  RNG: ``uniforms`` (explicit stream, restated ``choice``) if given,
    else ``rng.choice`` if given, else the process-global ``np.random.choice`` like the reference.
    ``s_override`` replaces the torch-CPU f32 weight sum (box/thread-count dependent)
    by a recorded one, so golden vectors made on another box replay exactly."""
    H, W = best_cert.shape
    if no_filter:                                   # core/sampling.py:15-21
        flat = torch.clamp(best_cert.clone().to("cpu"), max=cap).reshape(-1).numpy()
        if flat.size == 0:
            return np.zeros((0,), dtype=np.int64)
        return np.argsort(-flat)[:min(M, flat.size)]

    w, s = sample_weights(best_cert, cap, border)
    if s_override is not None:
        s = torch.tensor(np.float32(s_override))
    if taps is not None:
        taps["weights"] = w.numpy().copy()
        taps["s"] = np.float32(s.item())
    if s <= 0:                                      # core/sampling.py:27-28
        return np.zeros((0,), dtype=np.int64)
    p = (w / s).numpy()                             # f32 division, core/sampling.py:29
    m_main = int(M * 0.85)
    size = min(m_main, p.size)
    if uniforms is not None:
        main, used, rounds = legacy_choice_no_replace(p, size, uniforms)
        if taps is not None:
            taps["uniforms_used"] = used
            taps["rounds"] = rounds
    elif rng is not None:
        main = rng.choice(p.size, size=size, replace=False, p=p)
    else:
        main = np.random.choice(p.size, size=size, replace=False, p=p)
    cov = coverage_picks(p, H, W, tiles, M - len(main))
    if taps is not None:
        taps["p"] = p
        taps["idx_main"] = np.asarray(main, dtype=np.int64)
        taps["idx_cov"] = cov
    return np.unique(np.concatenate([main, cov]))   # ascending, deduped (core/sampling.py:52)


# ----------------------------------------------------------------------------------------------
# two-view geometry  (core/geometry.py:53-141)
# ----------------------------------------------------------------------------------------------
def fundamental_matrix(K1, R1, t1, K2, R2, t2) -> np.ndarray:
    """F = K2^-T [t]x R K1^-1 with R = R2 R1^T, t = t2 - R t1, all f32 (core/geometry.py:53-55,122-130)."""
    Rrel = R2 @ R1.T
    trel = (t2 - Rrel @ t1).reshape(3)
    a, b, c = trel.flatten()
    tx = np.array([[0, -c, b], [c, 0, -a], [-b, a, 0]], dtype=np.float32)
    return np.linalg.inv(K2).T @ (tx @ Rrel) @ np.linalg.inv(K1)


def sampson_distance(F: np.ndarray, uv1: np.ndarray, uv2: np.ndarray) -> np.ndarray:
    """First-order geometric error; evaluated in float64 because the homogeneous 1-column is f64
    (core/geometry.py:133-141)."""
    n = uv1.shape[0]
    h1 = np.concatenate([uv1, np.ones((n, 1))], axis=1)
    h2 = np.concatenate([uv2, np.ones((n, 1))], axis=1)
    l2 = (F @ h1.T).T
    l1 = (F.T @ h2.T).T
    num = np.sum(h2 * l2, axis=1)
    den = l2[:, 0] ** 2 + l2[:, 1] ** 2 + l1[:, 0] ** 2 + l1[:, 1] ** 2 + 1e-12
    return (num ** 2) / den


def dlt_points(P1: np.ndarray, P2: np.ndarray, uv1: np.ndarray, uv2: np.ndarray) -> np.ndarray:
    """Homogeneous DLT, null vector by f32 LAPACK SVD, dehomogenised with the +1e-12 guard
    (core/geometry.py:58-87; N==1 un-batched at :77-82)."""
    n = uv1.shape[0]
    if n == 0:
        return np.zeros((0, 4), dtype=np.float32)
    A = np.empty((n, 4, 4), dtype=np.float32)
    A[:, 0, :] = uv1[:, 0:1] * P1[2] - P1[0]
    A[:, 1, :] = uv1[:, 1:2] * P1[2] - P1[1]
    A[:, 2, :] = uv2[:, 0:1] * P2[2] - P2[0]
    A[:, 3, :] = uv2[:, 1:2] * P2[2] - P2[1]
    if n == 1:
        v = np.linalg.svd(A[0])[2][-1]
        w = v[3] if abs(v[3]) > 1e-12 else 1e-12
        return (v / w)[None, :]
    v = np.linalg.svd(A)[2][:, -1, :]
    w = np.where(np.abs(v[:, 3:4]) < 1e-12, 1e-12, v[:, 3:4])
    return v / w


def reprojection_error(P: np.ndarray, X: np.ndarray, uv: np.ndarray) -> np.ndarray:
    """core/geometry.py:91-104 (sgemm projection, z clamped at 1e-12)."""
    q = X @ P.T
    z = np.maximum(q[:, 2], 1e-12)
    du = q[:, 0] / z - uv[:, 0]
    dv = q[:, 1] / z - uv[:, 1]
    return np.sqrt(du * du + dv * dv)


def in_front(P: np.ndarray, X: np.ndarray) -> np.ndarray:
    """core/geometry.py:107-110."""
    return (P @ X.T)[2, :] > 0.0


def parallax_ok(C1: np.ndarray, C2: np.ndarray, X: np.ndarray, min_deg: float) -> np.ndarray:
    """core/geometry.py:113-119 (f32 throughout)."""
    r1 = X[:, :3] - C1.reshape(1, 3)
    r2 = X[:, :3] - C2.reshape(1, 3)
    r1 /= np.linalg.norm(r1, axis=1, keepdims=True) + 1e-12
    r2 /= np.linalg.norm(r2, axis=1, keepdims=True) + 1e-12
    ang = np.degrees(np.arccos(np.clip(np.sum(r1 * r2, axis=1), -1.0, 1.0)))
    return ang >= float(min_deg)


# ----------------------------------------------------------------------------------------------
# certainty post-processing between the matcher and the path  (core/pipeline.py:405-430)
# ----------------------------------------------------------------------------------------------
def resize_mask_nearest(mask_np: np.ndarray, H: int, W: int) -> np.ndarray:
    """``_mask_tensor_for_hw`` (core/pipeline.py:361-382): mask -> float32 [H, W].  ``F.interpolate(mode="nearest")``
    picks source index min(floor(dst * f32(in / out)), in - 1) per axis (ATen UpSample.h nearest_idx); equal
    shapes are passed through."""
    m = np.asarray(mask_np).astype(np.float32, copy=False)
    h, w = m.shape
    if (h, w) == (H, W):
        return m
    sy, sx = np.float32(h) / np.float32(H), np.float32(w) / np.float32(W)
    iy = np.minimum(np.floor(np.arange(H, dtype=np.float32) * sy).astype(np.int64), h - 1)
    ix = np.minimum(np.floor(np.arange(W, dtype=np.float32) * sx).astype(np.int64), w - 1)
    return m[iy[:, None], ix[None, :]]


def warp_mask_nearest(mask_hw: np.ndarray, grid_xy: np.ndarray) -> np.ndarray:
    """``F.grid_sample(mask, grid, mode="nearest", padding_mode="zeros", align_corners=False)`` (core/pipeline.py:423-429)
    for a single-channel [H, W] f32 mask and a [H, W, 2] f32 grid of normalised (x, y): un-normalise
    (g + 1) * (size / 2) - 0.5 in f32, round half to even, zero outside the image (NaN coordinates fall outside)."""
    H, W = mask_hw.shape
    gx = grid_xy[..., 0].astype(np.float32, copy=False)
    gy = grid_xy[..., 1].astype(np.float32, copy=False)
    with np.errstate(invalid="ignore", over="ignore"):
        fx = np.rint(((gx + np.float32(1)) * np.float32(W / 2)).astype(np.float32) - np.float32(0.5))
        fy = np.rint(((gy + np.float32(1)) * np.float32(H / 2)).astype(np.float32) - np.float32(0.5))
        ok = (fx > -1) & (fx < W) & (fy > -1) & (fy < H)
    ix = np.where(ok, fx, 0).astype(np.int64)
    iy = np.where(ok, fy, 0).astype(np.int64)
    return np.where(ok, mask_hw[iy, ix], np.float32(0)).astype(np.float32)


def certainty_prologue(cert_hw: np.ndarray, warp_hw: np.ndarray, maskA_np: Optional[np.ndarray],
                       maskB_np: Optional[np.ndarray], certainty_thresh: float) -> np.ndarray:
    """What ``_collect_reference_matches`` does to one pair's certainty map before the path sees it
    (core/pipeline.py:405-430), with every index made explicit: floor-clamp (NaN stays NaN), x maskA resized to the
    map, x maskB resized to the map and sampled at the warp's (xB, yB)."""
    c = np.asarray(cert_hw, dtype=np.float32)
    H, W = c.shape
    with np.errstate(invalid="ignore"):
        c = np.where(c < np.float32(certainty_thresh), np.float32(certainty_thresh), c).astype(np.float32)   # torch.clamp(min=)
        if maskA_np is not None:
            c = c * resize_mask_nearest(maskA_np, H, W)
        if maskB_np is not None:
            c = c * warp_mask_nearest(resize_mask_nearest(maskB_np, H, W), np.asarray(warp_hw)[..., 2:4])
    return c.astype(np.float32)


def certainty_prologue_torch(cert_hw: torch.Tensor, warp_hw: torch.Tensor, maskA_np: Optional[np.ndarray],
                             maskB_np: Optional[np.ndarray], certainty_thresh: float) -> torch.Tensor:
    """The same step written with the torch calls the reference makes (core/pipeline.py:405-430): pins the explicit
    index arithmetic of ``certainty_prologue`` to ATen's on this box."""
    import torch.nn.functional as F

    def to_hw(mask_np, hw):
        m = torch.from_numpy(np.asarray(mask_np).astype(np.float32, copy=False))
        if tuple(m.shape) != tuple(hw):
            m = F.interpolate(m.view(1, 1, m.shape[0], m.shape[1]), size=hw, mode="nearest").squeeze(0).squeeze(0)
        return m

    c = torch.clamp(torch.as_tensor(cert_hw), min=certainty_thresh)
    hw = (int(c.shape[0]), int(c.shape[1]))
    if maskA_np is not None:
        c = c * to_hw(maskA_np, hw)
    if maskB_np is not None:
        mb = to_hw(maskB_np, hw).view(1, 1, hw[0], hw[1])
        c = c * F.grid_sample(mb, torch.as_tensor(warp_hw)[..., 2:4].unsqueeze(0), mode="nearest",
                              padding_mode="zeros", align_corners=False).squeeze(0).squeeze(0)
    return c


# ----------------------------------------------------------------------------------------------
# per-reference driver  (core/pipeline.py:602-780)
# ----------------------------------------------------------------------------------------------
def best_neighbour(cert_list: Sequence[torch.Tensor]) -> Tuple[torch.Tensor, torch.Tensor]:
    """Per-pixel max over neighbours, lowest neighbour index on ties (core/pipeline.py:634-635)."""
    stack = torch.stack([torch.as_tensor(c) for c in cert_list], dim=0).to("cpu")
    return torch.max(stack, dim=0)


def decode_samples(warp_list, best_k: torch.Tensor, sel_idx: np.ndarray, w_match: int, h_match: int):
    """Winning-neighbour warp rows at the sampled pixels -> match-res pixel coords
    (core/pipeline.py:636-640,652-659).  Only the sampled rows are gathered (the reference gathers
    all H*W rows then indexes; same values)."""
    H, W = best_k.shape
    k = best_k.reshape(-1).numpy()[sel_idx]
    stack = torch.stack([torch.as_tensor(x) for x in warp_list], dim=0).to("cpu")
    rows = stack.reshape(stack.shape[0], H * W, 4)[torch.from_numpy(k), torch.from_numpy(sel_idx)].numpy()
    xA = (rows[:, 0] + 1.0) * 0.5 * (w_match - 1)
    yA = (rows[:, 1] + 1.0) * 0.5 * (h_match - 1)
    return k, xA, yA, rows[:, 2], rows[:, 3]


def bilinear_colour(img_u8: np.ndarray, xA: np.ndarray, yA: np.ndarray, w_match: int, h_match: int) -> np.ndarray:
    """core/pipeline.py:661-679: clipped-corner bilinear tap of the (resized) reference image; the
    int32 - float32 subtractions promote the weights to float64."""
    hh, ww = img_u8.shape[0], img_u8.shape[1]
    fx = xA * (ww / float(w_match))
    fy = yA * (hh / float(h_match))
    x0 = np.clip(np.floor(fx).astype(np.int32), 0, ww - 1)
    y0 = np.clip(np.floor(fy).astype(np.int32), 0, hh - 1)
    x1 = np.clip(x0 + 1, 0, ww - 1)
    y1 = np.clip(y0 + 1, 0, hh - 1)
    w00 = (x1 - fx) * (y1 - fy)
    w01 = (fx - x0) * (y1 - fy)
    w10 = (x1 - fx) * (fy - y0)
    w11 = (fx - x0) * (fy - y0)
    t00 = img_u8[y0, x0].astype(np.float32)
    t01 = img_u8[y0, x1].astype(np.float32)
    t10 = img_u8[y1, x0].astype(np.float32)
    t11 = img_u8[y1, x1].astype(np.float32)
    return (t00 * w00[:, None] + t01 * w01[:, None] + t10 * w10[:, None] + t11 * w11[:, None]) / 255.0


@dataclass
class OracleResult:
    xyz: np.ndarray
    rgb: np.ndarray
    err: np.ndarray
    sel_idx: np.ndarray
    debug_matches_by_nbr: Dict[int, np.ndarray] = field(default_factory=dict)
    debug_cert_by_nbr: Dict[int, np.ndarray] = field(default_factory=dict)
    taps: dict = field(default_factory=dict)


def triangulate_ref(cert_list, warp_list, img_u8: np.ndarray, ref_cam: OracleCamera,
                    nbr_cams: Sequence[OracleCamera], cfg: OracleConfig,
                    rng: Optional[np.random.RandomState] = None, uniforms: Optional[np.ndarray] = None,
                    collect_debug: bool = False, keep_taps: bool = False,
                    s_override: Optional[np.float32] = None) -> Optional[OracleResult]:
    """core/pipeline.py:602-780 for one reference view.  ``nbr_cams[k]`` is the camera of
    ``cert_list[k]`` (the reference's ``nn_ids[k]``).  Returns None where the reference does."""
    taps: dict = {}
    best_cert, best_k = best_neighbour(cert_list)
    H, W = best_cert.shape
    sel_idx = select_samples(best_cert, cfg.matches_per_ref, cap=cfg.sample_cap, border=cfg.border,
                             tiles=cfg.tiles, no_filter=cfg.no_filter, rng=rng, uniforms=uniforms,
                             taps=taps if keep_taps else None, s_override=s_override)
    if sel_idx.size == 0:
        return None
    wm, hm = cfg.w_match, cfg.h_match
    k_sel, xA, yA, xBn, yBn = decode_samples(warp_list, best_k, sel_idx, wm, hm)
    cert_sel = best_cert.reshape(-1).numpy()[sel_idx]
    rgb_all = bilinear_colour(img_u8, xA, yA, wm, hm)
    uvA_all = np.stack([xA * (ref_cam.width / float(wm)), yA * (ref_cam.height / float(hm))], axis=1)

    members: Dict[int, List[int]] = {}               # first-appearance order (core/pipeline.py:685-688)
    cam_of: Dict[int, OracleCamera] = {}
    for pos, kk in enumerate(k_sel):
        cam = nbr_cams[int(kk)]
        members.setdefault(cam.uid, []).append(pos)
        cam_of[cam.uid] = cam

    out_xyz, out_rgb, out_err = [], [], []
    dbg_m: Dict[int, np.ndarray] = {}
    dbg_c: Dict[int, np.ndarray] = {}
    denom = float(cfg.sample_cap) if float(cfg.sample_cap) > 1e-6 else 1.0
    group_taps = []
    for uid, pos_list in members.items():
        pos = np.asarray(pos_list, dtype=np.int64)
        cam = cam_of[uid]
        xB = (xBn[pos] + 1.0) * 0.5 * (wm - 1)
        yB = (yBn[pos] + 1.0) * 0.5 * (hm - 1)
        uvB = np.stack([xB * (cam.width / float(wm)), yB * (cam.height / float(hm))], axis=1)
        xAg, yAg, cg = xA[pos], yA[pos], cert_sel[pos]
        gt = {"uid": uid, "pos_all": pos.copy()}
        if (not cfg.no_filter) and cfg.sampson_thresh > 0:       # core/pipeline.py:708-727
            F = fundamental_matrix(ref_cam.K, ref_cam.R, ref_cam.t, cam.K, cam.R, cam.t)
            se = sampson_distance(F, uvA_all[pos], uvB)
            ok = se < float(cfg.sampson_thresh)
            gt["F"], gt["sampson"] = F, se
            if not np.any(ok):
                group_taps.append(gt)
                continue
            pos, xB, yB, uvB, xAg, yAg, cg = pos[ok], xB[ok], yB[ok], uvB[ok], xAg[ok], yAg[ok], cg[ok]
        if pos.size == 0:
            group_taps.append(gt)
            continue
        uvA = uvA_all[pos]
        X = dlt_points(ref_cam.P, cam.P, uvA, uvB)
        e = np.maximum(reprojection_error(ref_cam.P, X, uvA), reprojection_error(cam.P, X, uvB))
        if cfg.no_filter:                                        # core/pipeline.py:739-743
            keep = np.isfinite(X).all(axis=1) & np.isfinite(e)
        else:                                                    # core/pipeline.py:745-749
            keep = e <= float(cfg.reproj_thresh)
            keep &= in_front(ref_cam.P, X)
            keep &= in_front(cam.P, X)
            if cfg.min_parallax_deg > 0:
                keep &= parallax_ok(ref_cam.C, cam.C, X, cfg.min_parallax_deg)
        gt.update(pos=pos, X=X, err=e, keep=keep, uvA=uvA, uvB=uvB, P1=ref_cam.P, P2=cam.P)
        group_taps.append(gt)
        if not np.any(keep):
            continue
        out_xyz.append(X[keep][:, :3].astype(np.float32))
        out_rgb.append(rgb_all[pos][keep].astype(np.float32))
        out_err.append(e[keep].astype(np.float32))
        if collect_debug:                                        # core/pipeline.py:761-769
            m = np.stack([np.clip(xAg[keep], 0.0, float(wm - 1)), np.clip(yAg[keep], 0.0, float(hm - 1)),
                          np.clip(xB[keep], 0.0, float(wm - 1)), np.clip(yB[keep], 0.0, float(hm - 1))], axis=1)
            dbg_m[uid] = m.astype(np.float32, copy=False)
            dbg_c[uid] = np.clip(cg[keep] / denom, 0.0, 1.0).astype(np.float32, copy=False)
    if not out_xyz:
        return None
    if keep_taps:
        taps.update(ref_uid=ref_cam.uid, best_k=best_k.numpy(), best_cert=best_cert.numpy(), k_sel=k_sel, xA=xA, yA=yA,
                    xBn=xBn, yBn=yBn, uvA_all=uvA_all, rgb_all=rgb_all, groups=group_taps)
    return OracleResult(xyz=np.concatenate(out_xyz, axis=0), rgb=np.concatenate(out_rgb, axis=0),
                        err=np.concatenate(out_err, axis=0), sel_idx=sel_idx,
                        debug_matches_by_nbr=dbg_m, debug_cert_by_nbr=dbg_c, taps=taps)


# ----------------------------------------------------------------------------------------------
# output contract  (core/image_utils.py:24-26, core/writers.py:15-46)
# ----------------------------------------------------------------------------------------------
def to_uint8_rgb(rgb01: np.ndarray) -> np.ndarray:
    """clip(round(x*255)) with numpy's round-half-to-even (core/image_utils.py:24-26)."""
    return np.clip(np.round(rgb01 * 255.0), 0, 255).astype(np.uint8)


def ply_bytes(xyz: np.ndarray, rgb_u8: np.ndarray) -> bytes:
    """The exact byte stream core/writers.py:29-46 writes (binary little-endian PLY)."""
    import struct
    n = xyz.shape[0]
    head = ("ply\nformat binary_little_endian 1.0\n" f"element vertex {n}\n"
            "property float x\nproperty float y\nproperty float z\n"
            "property uchar red\nproperty uchar green\nproperty uchar blue\nend_header\n")
    parts = [head.encode("ascii")]
    for i in range(n):
        parts.append(struct.pack("<fff", float(xyz[i, 0]), float(xyz[i, 1]), float(xyz[i, 2])))
        parts.append(struct.pack("BBB", int(rgb_u8[i, 0]), int(rgb_u8[i, 1]), int(rgb_u8[i, 2])))
    return b"".join(parts)


def points3d_bin_bytes(xyz: np.ndarray, rgb_u8: np.ndarray, errors: Optional[np.ndarray] = None) -> bytes:
    """The exact byte stream core/writers.py:15-26 writes (COLMAP-like points3D.bin without tracks)."""
    import struct
    n = xyz.shape[0]
    if errors is None:
        errors = np.zeros((n,), dtype=np.float32)
    parts = [struct.pack("<Q", n)]
    for i in range(n):
        parts.append(struct.pack("<Q", i + 1))
        parts.append(struct.pack("<ddd", float(xyz[i, 0]), float(xyz[i, 1]), float(xyz[i, 2])))
        parts.append(struct.pack("<BBB", int(rgb_u8[i, 0]), int(rgb_u8[i, 1]), int(rgb_u8[i, 2])))
        parts.append(struct.pack("<d", float(errors[i])))
    return b"".join(parts)


# ---------------------------------------------------------------------------------------------
# Reducers after the path -- pinned by tests/golden/output_reducers.npz, which
# tests/golden/make_output_golden.py froze from the live reference functions.
# ---------------------------------------------------------------------------------------------
def apply_point_cap(xyz: np.ndarray, rgb: np.ndarray, err: np.ndarray, max_points: int, seed: int):
    """densify.py:110-120: a PCG64 ``default_rng(seed).choice`` without replacement, then three gathers."""
    if max_points > 0 and xyz.shape[0] > max_points:
        sel = np.random.default_rng(seed).choice(xyz.shape[0], size=max_points, replace=False)
        return xyz[sel], rgb[sel], err[sel]
    return xyz, rgb, err


def preview_subsample(matches: np.ndarray, cert_norm: np.ndarray, ref_id: int, nbr_id: int, max_matches: int = 10000):
    """core/pipeline.py:569-582: at most ``max_matches`` of a pair's kept matches, seeded by the pair's ids."""
    matches = np.asarray(matches, dtype=np.float32)
    cert_norm = np.asarray(cert_norm, dtype=np.float32)
    if matches.shape[0] > max_matches > 0:
        seed = ((int(ref_id) & 0xFFFF_FFFF) * 73856093) ^ ((int(nbr_id) & 0xFFFF_FFFF) * 19349663)
        rng = np.random.default_rng(seed & 0xFFFF_FFFF)
        sel_idx = rng.choice(matches.shape[0], size=max_matches, replace=False)
        matches = matches[sel_idx]
        cert_norm = cert_norm[sel_idx]
    return matches, cert_norm


# ---------------------------------------------------------------------------------------------
# Pair generation -- pinned by tests/golden/selection.npz (tests/golden/make_selection_golden.py
# froze the live reference's select_cameras_kcenters / nearest_neighbors on synthetic camera sets).
# ---------------------------------------------------------------------------------------------
def _row_sum16_f32(a: np.ndarray) -> np.ndarray:
    """np.add.reduce over 16 contiguous float32 values, as numpy evaluates it: 8 running sums, then a fixed tree
    (measured: equals np.linalg.norm(.., axis=1) ** 2 bit for bit)."""
    r = a[:, :8] + a[:, 8:16]
    return ((r[:, 0] + r[:, 1]) + (r[:, 2] + r[:, 3])) + ((r[:, 4] + r[:, 5]) + (r[:, 6] + r[:, 7]))


def _einsum_row_sum16_f32(P: np.ndarray) -> np.ndarray:
    """Row sums of 16 float32 products the way np.einsum("nd,nd->n") forms them in numpy's x86-64 wheels
    (numpy/_core/src/multiarray/einsum_sumprod.c.src, sum_of_products_contig_contig_outstride0_two: the einsum loops are built for
    the baseline SIMD width - 4 lanes, multiply then add, no FMA - and 16 elements are one pass of the 4x unrolled loop, whose
    accumulator chain starts at the LAST vector): lane j adds p[12+j], p[8+j], p[4+j], p[j] in that order, then
    (l0 + l1) + (l2 + l3).  Measured against np.einsum on 20 000 random rows (tests/test_oracle_selection.py)."""
    f32 = np.float32
    P = np.asarray(P, dtype=f32)
    lanes = P[:, 12:16].copy()
    for b in (8, 4, 0):
        lanes = (lanes + P[:, b:b + 4]).astype(f32)
    return ((lanes[:, 0] + lanes[:, 1]).astype(f32) + (lanes[:, 2] + lanes[:, 3]).astype(f32)).astype(f32)


def select_cameras_kcenters(flat_poses: np.ndarray, k: int, explicit: bool = True):
    """core/selection.py:36-54 with the float32 operation order written out (the order the CUDA kernel mirrors):
    column mean / std accumulate row by row; the first pick is the arg-max of np.einsum's row reduction (its own order:
    ``_einsum_row_sum16_f32`` - it decides ties between rows of equal norm, as on symmetric rings); the distances'
    row norms are np.linalg.norm's 8-accumulator pairwise sum.  Returns (sorted centres, pick order)."""
    f32 = np.float32
    X = np.asarray(flat_poses, dtype=f32)
    n = X.shape[0]
    k = max(1, min(int(k), n))
    s = np.zeros(X.shape[1], f32)
    for i in range(n):
        s = s + X[i]
    mu = s / f32(n)
    d = X - mu
    d2 = d * d
    s2 = np.zeros(X.shape[1], f32)
    for i in range(n):
        s2 = s2 + d2[i]
    sigma = np.sqrt(s2 / f32(n)) + f32(1e-8)
    Xn = (X - mu) / sigma
    first = int(np.argmax(_einsum_row_sum16_f32(Xn * Xn)))
    centers = [first]
    diff = Xn - Xn[first]
    dist = np.sqrt(_row_sum16_f32(diff * diff))
    dist[first] = -np.inf
    for _ in range(1, k):
        nxt = int(np.argmax(dist))
        centers.append(nxt)
        diff = Xn - Xn[nxt]
        dist = np.minimum(dist, np.sqrt(_row_sum16_f32(diff * diff)))
        dist[nxt] = -np.inf
    return sorted(centers), centers


def nearest_neighbors_exact(flat_poses: np.ndarray, k: int):
    """core/selection.py:57-70 with distances formed from differences in float64 (what torch.cdist approximates with one
    float32 sgemm): indices [n, k] ascending by (distance, index), and the distances."""
    X = np.asarray(flat_poses, dtype=np.float32).astype(np.float64)
    n = X.shape[0]
    if n <= 1:
        return np.empty((n, 0), dtype=np.int64), np.empty((n, 0))
    k = max(1, min(int(k), n - 1))
    D = np.sqrt(((X[:, None, :] - X[None, :, :]) ** 2).sum(-1))
    np.fill_diagonal(D, np.inf)
    idx = np.argsort(D, axis=1, kind="stable")[:, :k]
    return idx.astype(np.int64), np.take_along_axis(D, idx, axis=1)


def cdist_squared_f32(flat_poses: np.ndarray) -> np.ndarray:
    """The float32 matrix torch.cdist takes the square root of (core/selection.py:66), restated operation by operation
    (ATen ``_euclidean_dist``, used for more than 25 rows): ``x1_ = [-2 x, |x|^2, 1]``, ``x2_ = [x, 1, |x|^2]``,
    ``x1_ @ x2_.T`` - one sgemm with K = 18, whose K loop is a chain of float32 FMAs in index order, and row norms summed as
    eight lanes (a[i] + a[i + 8]) added left to right.  Both orders were measured against torch 2.11 on the build box
    (bit-identical on rings of 40 / 185 / 1000 views and on random poses, 1 and 8 threads)."""
    X = np.asarray(flat_poses, dtype=np.float32)
    n, K = X.shape
    sq = (X * X).astype(np.float32)
    if K == 16:
        lanes = (sq[:, :8] + sq[:, 8:]).astype(np.float32)
        nrm = lanes[:, 0].copy()
        for i in range(1, 8):
            nrm = (nrm + lanes[:, i]).astype(np.float32)
    else:
        nrm = sq.sum(axis=1, dtype=np.float32)
    a = np.concatenate([X * np.float32(-2.0), nrm[:, None], np.ones((n, 1), np.float32)], axis=1)
    b = np.concatenate([X, np.ones((n, 1), np.float32), nrm[:, None]], axis=1)
    acc = np.zeros((n, n), dtype=np.float32)
    for k in range(a.shape[1]):              # fma(a, b, acc): the product of two float32 is exact in float64
        acc = (acc.astype(np.float64) + a[:, k:k + 1].astype(np.float64) * b[None, :, k].astype(np.float64)).astype(np.float32)
    return acc


def cdist_f32(flat_poses: np.ndarray) -> np.ndarray:
    """torch.cdist(x, x, p=2) for float32 rows, bit for bit.  ATen picks the matrix-multiply formulation above only when there are
    more than 25 rows (cdist_impl: ``r1 > 25 || r2 > 25``); up to 25 rows - small scenes - it runs the direct kernel
    (aten/src/ATen/native/cpu/DistanceOpsKernel.cpp, run_parallel_cdist with tdist_calc): per pair a sequential float32 sum over
    the columns of (a - b) * (a - b), multiply then add (no FMA), then sqrt.  Both measured against torch 2.11 here."""
    X = np.asarray(flat_poses, dtype=np.float32)
    n = X.shape[0]
    if n > 25:
        return np.sqrt(np.maximum(cdist_squared_f32(X), np.float32(0.0)))
    agg = np.zeros((n, n), dtype=np.float32)
    for k in range(X.shape[1]):
        d = np.abs((X[:, None, k] - X[None, :, k]).astype(np.float32))
        agg = (agg + (d * d).astype(np.float32)).astype(np.float32)
    return np.sqrt(agg)


# ---- torch.topk(largest=False) on a CPU row, including the order it leaves among EXACTLY equal values -------------------
# ATen's CPU kernel (aten/src/ATen/native/cpu/SortingKernel.cpp, topk_impl_loop; not in /root/reference: torch 2.11 is the
# reference's dependency) copies the row into (value, index) pairs and runs, with a comparator that looks at the value only
# (NaN last),   k * 64 <= n :  std::partial_sort(begin, begin + k, end)
#               otherwise   :  std::nth_element(begin, begin + k - 1, end) ; std::sort(begin, begin + k - 1)
# Both are deterministic algorithms; which of several equal values end up in the first k places, and in what order, is decided
# by their moves.  libstdc++'s versions (bits/stl_algo.h, bits/stl_heap.h: heap select + sort_heap; introselect with the
# median-of-three pivot moved to the front, unguarded Hoare partition, insertion sort below four elements; introsort with a
# threshold of 16) are restated below move for move.  Pinned against torch.topk itself on tie-heavy random rows
# (tests/test_oracle_selection.py) and through the frozen neighbour tables of the live reference.

def _topk_lt(x, y) -> bool:
    a, b = x[0], y[0]
    return ((a == a) and (b != b)) or a < b


def _stl_push_heap(v, first, hole, top, value):
    parent = (hole - 1) // 2
    while hole > top and _topk_lt(v[first + parent], value):
        v[first + hole] = v[first + parent]
        hole = parent
        parent = (hole - 1) // 2
    v[first + hole] = value


def _stl_adjust_heap(v, first, hole, length, value):
    top = hole
    child = hole
    while child < (length - 1) // 2:
        child = 2 * (child + 1)
        if _topk_lt(v[first + child], v[first + child - 1]):
            child -= 1
        v[first + hole] = v[first + child]
        hole = child
    if (length & 1) == 0 and child == (length - 2) // 2:
        child = 2 * (child + 1)
        v[first + hole] = v[first + child - 1]
        hole = child - 1
    _stl_push_heap(v, first, hole, top, value)


def _stl_make_heap(v, first, last):
    length = last - first
    if length < 2:
        return
    parent = (length - 2) // 2
    while True:
        _stl_adjust_heap(v, first, parent, length, v[first + parent])
        if parent == 0:
            return
        parent -= 1


def _stl_pop_heap(v, first, last, result):
    value = v[result]
    v[result] = v[first]
    _stl_adjust_heap(v, first, 0, last - first, value)


def _stl_heap_select(v, first, middle, last):
    _stl_make_heap(v, first, middle)
    for i in range(middle, last):
        if _topk_lt(v[i], v[first]):
            _stl_pop_heap(v, first, middle, i)


def _stl_partial_sort(v, first, middle, last):
    _stl_heap_select(v, first, middle, last)
    while middle - first > 1:                       # std::__sort_heap
        middle -= 1
        _stl_pop_heap(v, first, middle, middle)


def _stl_unguarded_partition_pivot(v, first, last):
    mid = first + (last - first) // 2
    a, b, c = first + 1, mid, last - 1              # std::__move_median_to_first(first, first + 1, mid, last - 1)
    if _topk_lt(v[a], v[b]):
        m = b if _topk_lt(v[b], v[c]) else (c if _topk_lt(v[a], v[c]) else a)
    else:
        m = a if _topk_lt(v[a], v[c]) else (c if _topk_lt(v[b], v[c]) else b)
    v[first], v[m] = v[m], v[first]
    lo, hi = first + 1, last                        # std::__unguarded_partition(first + 1, last, pivot = first)
    while True:
        while _topk_lt(v[lo], v[first]):
            lo += 1
        hi -= 1
        while _topk_lt(v[first], v[hi]):
            hi -= 1
        if not lo < hi:
            return lo
        v[lo], v[hi] = v[hi], v[lo]
        lo += 1


def _stl_unguarded_linear_insert(v, last):
    val = v[last]
    nxt = last - 1
    while _topk_lt(val, v[nxt]):
        v[last] = v[nxt]
        last = nxt
        nxt -= 1
    v[last] = val


def _stl_insertion_sort(v, first, last):
    for i in range(first + 1, last):
        if _topk_lt(v[i], v[first]):
            val = v[i]
            v[first + 1:i + 1] = v[first:i]         # std::move_backward
            v[first] = val
        else:
            _stl_unguarded_linear_insert(v, i)


def _stl_nth_element(v, first, nth, last):
    if first == last or nth == last:
        return
    depth = ((last - first).bit_length() - 1) * 2
    while last - first > 3:
        if depth == 0:
            _stl_heap_select(v, first, nth + 1, last)
            v[first], v[nth] = v[nth], v[first]
            return
        depth -= 1
        cut = _stl_unguarded_partition_pivot(v, first, last)
        if cut <= nth:
            first = cut
        else:
            last = cut
    _stl_insertion_sort(v, first, last)


def _stl_sort(v, first, last):
    def loop(first, last, depth):
        while last - first > 16:
            if depth == 0:
                _stl_partial_sort(v, first, last, last)
                return
            depth -= 1
            cut = _stl_unguarded_partition_pivot(v, first, last)
            loop(cut, last, depth)
            last = cut
    if first == last:
        return
    loop(first, last, ((last - first).bit_length() - 1) * 2)
    if last - first > 16:
        _stl_insertion_sort(v, first, first + 16)
        for i in range(first + 16, last):
            _stl_unguarded_linear_insert(v, i)
    else:
        _stl_insertion_sort(v, first, last)


def topk_smallest_like_torch(row, k: int) -> list:
    """Indices ``torch.topk(row, k, largest=False)`` returns for a CPU row (see the block comment above)."""
    n = len(row)
    v = [(float(row[j]), j) for j in range(n)]
    if k * 64 <= n:
        _stl_partial_sort(v, 0, k, n)
    else:
        _stl_nth_element(v, 0, k - 1, n)
        _stl_sort(v, 0, k - 1)
    return [v[j][1] for j in range(k)]


def nearest_neighbors_cdist(flat_poses: np.ndarray, k: int) -> np.ndarray:
    """core/selection.py:57-70 with torch.cdist's own float32 arithmetic (``cdist_squared_f32``) and torch.topk's own
    selection (``topk_smallest_like_torch``): the reference's neighbour table, index for index.  Ring cameras have left /
    right neighbours at mathematically equal distance; which one comes first is decided by the rounding of cdist's formula
    and, where the float32 distances are exactly equal, by the moves of std::partial_sort / std::nth_element."""
    n = int(np.asarray(flat_poses).shape[0])
    if n <= 1:
        return np.empty((n, 0), dtype=np.int64)
    k = max(1, min(int(k), n - 1))
    D = cdist_f32(flat_poses)
    np.fill_diagonal(D, np.inf)
    return np.asarray([topk_smallest_like_torch(D[i], k) for i in range(n)], dtype=np.int64)


def voxel_downsample(xyz: np.ndarray, rgb: np.ndarray, voxel_size: float):
    """densify.py:29-50 with Open3D's voxel_down_sample restated from its published source
    (open3d/geometry/PointCloud.cpp: VoxelDownSample, AccumulatedPoint).  PARITY UNPINNED: Open3D is neither in the
    reference tree nor installed here, so this restatement could not be run against it.  Voxels are returned in the order
    of their first point (Open3D: std::unordered_map iteration order, implementation-defined)."""
    pts = np.asarray(xyz).astype(np.float64)
    rgb = np.asarray(rgb)
    cols = rgb[:, :3].astype(np.float64) / 255.0 if rgb.max() > 1.0 else rgb[:, :3].astype(np.float64)      # :41-44
    vmin = pts.min(axis=0) - float(voxel_size) * 0.5
    idx = np.floor((pts - vmin) / float(voxel_size)).astype(np.int64)
    uniq, first, inv = np.unique(idx, axis=0, return_index=True, return_inverse=True)
    inv = np.asarray(inv).reshape(-1)
    rank = np.empty(len(uniq), dtype=np.int64)
    rank[np.argsort(first, kind="stable")] = np.arange(len(uniq))
    v = rank[inv]
    acc_p = np.zeros((len(uniq), 3))
    acc_c = np.zeros((len(uniq), 3))
    np.add.at(acc_p, v, pts)            # unbuffered: accumulates in point order, like Open3D's loop
    np.add.at(acc_c, v, cols)
    cnt = np.bincount(v, minlength=len(uniq)).astype(np.float64)[:, None]
    return (acc_p / cnt).astype(np.float32), (acc_c / cnt).astype(np.float32)
