"""CPU: the oracle's pair-generation restatement against the live reference's frozen outputs (tests/golden/selection.npz)."""
import os

import numpy as np

from oracle import densify_oracle as O
from tests.golden.make_selection_golden import selection_cases
from tests.helpers import GOLDEN_DIR


def nn_equivalent(idx, idx_ref, flat, tol):
    """Same neighbour table up to the order / choice among views whose distances differ by <= tol."""
    X = np.asarray(flat, np.float32).astype(np.float64)
    if idx.shape != idx_ref.shape:
        return False
    for i in range(idx.shape[0]):
        a, b = idx[i], idx_ref[i]
        if len(set(a.tolist())) != a.size or i in a:
            return False
        da = np.sqrt(((X[a] - X[i]) ** 2).sum(-1))
        db = np.sqrt(((X[b] - X[i]) ** 2).sum(-1))
        if not np.all(np.abs(da - db) <= tol):
            return False
    return True


def tie_report(idx, idx_ref, D):
    """Rows that differ from the reference, and whether every difference is a permutation among views whose float32
    distances (``D``: torch.cdist's own values) are EXACTLY equal - the order torch.topk leaves to std::nth_element."""
    rows = np.nonzero((idx != idx_ref).any(axis=1))[0]
    only_ties = all(np.array_equal(D[i, idx[i]], D[i, idx_ref[i]]) for i in rows)      # same distances, place by place
    return rows, only_ties


def test_nearest_neighbours_mirror_cdist_and_differ_only_on_exact_ties():
    """The cdist-mirroring restatement gives the live reference's neighbour table except for the order inside groups of
    exactly equal float32 distances (ring cameras: left and right neighbour), which torch.topk does not define."""
    z = np.load(os.path.join(GOLDEN_DIR, "selection.npz"))
    for name, (flat, _, kn) in selection_cases().items():
        ref = z[f"{name}_nn"]
        if ref.shape[0] <= 1:
            continue
        idx = O.nearest_neighbors_cdist(flat, kn)
        D = np.sqrt(np.maximum(O.cdist_squared_f32(flat), np.float32(0.0)))
        rows, only_ties = tie_report(idx, ref, D)
        print(f"[knn] {name}: {len(rows)} of {ref.shape[0]} rows differ from the reference, all inside exact float32 ties: {only_ties}")
        assert only_ties, name
        # and the sorted distances of every row are the reference's, bit for bit
        assert np.array_equal(np.take_along_axis(D, idx, 1), np.take_along_axis(D, ref, 1)), name


def test_kcenters_explicit_order_equals_reference():
    z = np.load(os.path.join(GOLDEN_DIR, "selection.npz"))
    for name, (flat, kc, _) in selection_cases().items():
        got, order = O.select_cameras_kcenters(flat, kc)
        assert np.array_equal(np.asarray(got), z[f"{name}_centers"]), name
        assert sorted(order) == got and len(set(order)) == len(order)


def test_nearest_neighbours_equal_reference_up_to_cdist_noise():
    z = np.load(os.path.join(GOLDEN_DIR, "selection.npz"))
    for name, (flat, _, kn) in selection_cases().items():
        idx, _ = O.nearest_neighbors_exact(flat, kn)
        ref = z[f"{name}_nn"]
        assert idx.shape == ref.shape, name
        # torch.cdist's |x|^2 + |y|^2 - 2 x.y carries ~1e-3 of cancellation error: views closer than that swap places
        assert nn_equivalent(idx, ref, flat, 5e-3), name
        if name.startswith("random"):
            assert np.array_equal(idx, ref), name
