"""CPU: vectorised writers reproduce the reference's bytes (golden made by the live reference writers)."""
import os

import numpy as np

from lichtfeld_densification_plugin_b200.core import writers as W
from tests.helpers import GOLDEN_DIR


def test_writers_byte_identical_to_reference_golden(tmp_path):
    z = np.load(os.path.join(GOLDEN_DIR, "writers.npz"))
    u8 = W.to_uint8_rgb(z["rgb"])
    assert np.array_equal(u8, z["rgb_u8"])
    W.write_ply(str(tmp_path / "a.ply"), z["xyz"], u8)
    W.write_points3D_bin(str(tmp_path / "a.bin"), z["xyz"], u8, z["err"])
    W.write_points3D_bin(str(tmp_path / "b.bin"), z["xyz"], u8, None)
    assert (tmp_path / "a.ply").read_bytes() == z["ply"].tobytes()
    assert (tmp_path / "a.bin").read_bytes() == z["bin"].tobytes()
    assert (tmp_path / "b.bin").read_bytes() == z["bin_noerr"].tobytes()


def test_writers_empty_and_half_to_even(tmp_path):
    W.write_ply(str(tmp_path / "e.ply"), np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint8))
    assert (tmp_path / "e.ply").read_bytes() == W.ply_header(0)
    W.write_points3D_bin(str(tmp_path / "e.bin"), np.zeros((0, 3), np.float32), np.zeros((0, 3), np.uint8))
    assert (tmp_path / "e.bin").read_bytes() == (0).to_bytes(8, "little")
    x = np.array([0.5 / 255, 1.5 / 255, 2.5 / 255, -1.0, 2.0], dtype=np.float64)
    assert W.to_uint8_rgb(x).tolist() == [0, 2, 2, 0, 255]
