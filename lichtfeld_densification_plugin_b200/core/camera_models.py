"""Camera records and the packed per-launch camera tables.

``CameraRecord`` keeps the reference's field names and dtypes (reference core/camera_models.py:10-21:
K,R [3,3] f32; t [3,1] f32; P [3,4] f32; C [3] f32; width/height = full-resolution camera size).
Only pinhole intrinsics are honoured -- like the reference, which drops distortion coefficients
(reference core/geometry.py:10-30).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np


@dataclass
class CameraRecord:
    uid: int
    image_path: str
    width: int
    height: int
    K: np.ndarray
    R: np.ndarray
    t: np.ndarray
    P: np.ndarray
    C: np.ndarray
    mask_path: Optional[str] = None

    @classmethod
    def from_KRt(cls, uid: int, width: int, height: int, K, R, t, image_path: str = "",
                 mask_path: Optional[str] = None) -> "CameraRecord":
        """Build P and C from K,R,t with the reference's float32 arithmetic
        (reference densify.py:226-230, core/geometry.py:45-50)."""
        K = np.asarray(K, dtype=np.float32).reshape(3, 3)
        R = np.asarray(R, dtype=np.float32).reshape(3, 3)
        t = np.asarray(t, dtype=np.float32).reshape(3, 1)
        P = K @ np.concatenate([R, t], axis=1)
        C = (-R.T @ t).reshape(3)
        return cls(uid=int(uid), image_path=image_path, width=int(width), height=int(height),
                   K=K, R=R, t=t, P=P, C=C, mask_path=mask_path)

    def flat_pose(self) -> np.ndarray:
        """Row-major 4x4 world-to-camera transform, used for clustering / nearest neighbours."""
        T = np.eye(4)
        T[:3, :3] = self.R
        T[:3, 3] = self.t.reshape(3)
        return T.reshape(-1)


def stack_flat_poses(records: Sequence[CameraRecord]) -> np.ndarray:
    return np.stack([r.flat_pose() for r in records], axis=0)
