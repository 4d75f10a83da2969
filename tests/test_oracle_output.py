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
