"""CPU: the oracle's restatement of the post-path reducers equals the live reference's frozen outputs."""
import os

import numpy as np

from oracle import densify_oracle as O
from tests.golden.make_output_golden import CAP_CASES, PREVIEW_CASES, cap_inputs, preview_inputs
from tests.helpers import GOLDEN_DIR


def test_point_cap_matches_reference_golden():
    z = np.load(os.path.join(GOLDEN_DIR, "output_reducers.npz"))
    for i, c in enumerate(CAP_CASES):
        assert z[f"cap{i}_case"].tolist() == [c["n"], c["max_points"], c["seed"], c["data_seed"]]
        xyz, rgb, err = cap_inputs(c["n"], c["data_seed"])
        a, b, e = O.apply_point_cap(xyz, rgb, err, c["max_points"], c["seed"])
        assert np.array_equal(a, z[f"cap{i}_xyz"]) and np.array_equal(b, z[f"cap{i}_rgb"]) and np.array_equal(e, z[f"cap{i}_err"])


def test_preview_subsample_matches_reference_golden():
    z = np.load(os.path.join(GOLDEN_DIR, "output_reducers.npz"))
    for i, c in enumerate(PREVIEW_CASES):
        m, cn = preview_inputs(c["k"], c["data_seed"])
        a, b = O.preview_subsample(m, cn, c["ref_id"], c["nbr_id"], c["max_matches"])
        assert np.array_equal(a, z[f"pv{i}_matches"]) and np.array_equal(b, z[f"pv{i}_cert"])
        assert a.shape[0] == min(c["k"], c["max_matches"])


def test_voxel_downsample_restatement_properties():
    """PARITY UNPINNED (Open3D is not installed): properties the published algorithm implies."""
    rs = np.random.RandomState(1)
    xyz = (rs.standard_normal((3000, 3)) * 2).astype(np.float32)
    rgb = rs.random_sample((3000, 3)).astype(np.float32)
    vs = 0.3
    p, c = O.voxel_downsample(xyz, rgb, vs)
    vmin = xyz.astype(np.float64).min(0) - vs * 0.5
    idx_in = np.floor((xyz.astype(np.float64) - vmin) / vs).astype(np.int64)
    idx_out = np.floor((p.astype(np.float64) - vmin) / vs + 1e-9 * 0).astype(np.int64)
    assert p.shape[0] == len(np.unique(idx_in, axis=0)) and p.dtype == np.float32 and c.dtype == np.float32
    assert len(np.unique(idx_out, axis=0)) >= p.shape[0] - 3            # a mean may round onto a voxel face
    one_p, one_c = O.voxel_downsample(xyz, rgb * 255.0, 1e3)            # one voxel; colours given as 0..255
    assert one_p.shape == (1, 3)
    np.testing.assert_allclose(one_p[0], xyz.astype(np.float64).mean(0), rtol=1e-6)
    np.testing.assert_allclose(one_c[0], rgb.astype(np.float64).mean(0), rtol=1e-5)
    assert np.array_equal(O.voxel_downsample(xyz[:1], rgb[:1], vs)[0], xyz[:1])
