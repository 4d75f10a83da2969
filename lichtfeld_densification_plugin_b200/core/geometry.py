"""Host-side camera constants for the densification kernels.

Only the per-view / per-pair constants are computed on the host (a few 3x3 float32 products per
pair); all per-point geometry runs in csrc/ldp_geometry.cu.  The constants must be the reference's
float32 values bit for bit because the kernels reproduce the reference's float32 arithmetic from
them, so these helpers keep numpy float32 matmul / LAPACK inverse like the reference does
(reference core/geometry.py:45-55,122-130, densify.py:226-230).
"""
from __future__ import annotations

import numpy as np


def P_from_KRt(K: np.ndarray, R: np.ndarray, t: np.ndarray) -> np.ndarray:
    """3x4 projection K [R | t]."""
    return K @ np.hstack([R, t.reshape(3, 1)])


def cam_center_world(R: np.ndarray, t: np.ndarray) -> np.ndarray:
    """Camera centre C = -R^T t as a flat 3-vector."""
    return (-(R.T) @ t.reshape(3, 1)).reshape(3)


def skew(v: np.ndarray) -> np.ndarray:
    """Cross-product matrix [v]x in float32."""
    x, y, z = np.asarray(v).reshape(-1)[:3]
    out = np.zeros((3, 3), dtype=np.float32)
    out[0, 1], out[0, 2] = -z, y
    out[1, 0], out[1, 2] = z, -x
    out[2, 0], out[2, 1] = -y, x
    return out


def fundamental_from_world2cam(K1, R1, t1, K2, R2, t2) -> np.ndarray:
    """F with x2^T F x1 = 0 for world-to-camera poses (R_i, t_i): K2^-T [t]x R K1^-1,
    R = R2 R1^T, t = t2 - R t1 (reference core/geometry.py:122-130)."""
    R = R2 @ R1.T
    t = (t2 - R @ t1).reshape(3)
    E = skew(t) @ R
    return (np.linalg.inv(K2).T @ E) @ np.linalg.inv(K1)
