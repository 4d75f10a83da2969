"""Sampling strategies for certainty maps -- GPU version of the reference's ``core/sampling.py``.

``select_samples_with_coverage(cert_map, M, cap, border, tiles, no_filter)`` keeps the reference's signature and
semantics (reference core/sampling.py:8-53): 85 % certainty-weighted draws without replacement (numpy's legacy
``RandomState.choice`` algorithm on the process-global MT19937 stream) + per-tile best-pixel coverage picks,
returned sorted and unique; ``no_filter``: the M pixels of largest capped certainty in descending order.
The work runs in ``ldp_sample_refs`` (csrc/ldp_sample.cu); there is no CPU fallback.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import _native as N
from ..engine import PathConfig


def select_samples_with_coverage(cert_map: torch.Tensor, M: int, cap: float = 0.9, border: int = 2, tiles: int = 24,
                                 no_filter: bool = False) -> np.ndarray:
    from .pipeline import _mt_stream_from_global, _raise_for_status, get_engine
    eng = get_engine()
    cert = cert_map.detach().to(eng.device, torch.float32).contiguous()
    H, W = int(cert.shape[0]), int(cert.shape[1])
    if H * W == 0:
        return np.zeros((0,), dtype=np.int64)
    cfg = PathConfig(matches_per_ref=int(M), sample_cap=float(cap), border=int(border), tiles=int(tiles), no_filter=bool(no_filter))
    n_uni = 2 * int(M * 0.85) + 64
    while True:
        batch = eng.new_batch(H, W, W, H)
        batch.add_cert_only([cert])
        u = None
        if not no_filter:
            u = torch.from_numpy(_mt_stream_from_global(n_uni)[None, :]).to(eng.device)
        sel, n_samples, status, used = eng.sample(batch, cfg, u)
        code = int(status[0].item()) & N.LDP_REF_CODE_MASK
        if code == N.LDP_REF_UNIFORMS_EXHAUSTED:
            n_uni *= 2
            continue
        break
    if code == N.LDP_REF_EMPTY:
        return np.zeros((0,), dtype=np.int64)
    _raise_for_status(code)
    if int(used[0].item()) > 0:
        np.random.random_sample(int(used[0].item()))        # advance the global stream like np.random.choice did
    return sel[0, : int(n_samples[0].item())].cpu().numpy().astype(np.int64)
