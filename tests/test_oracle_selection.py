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


def test_nearest_neighbours_equal_the_reference_table_index_for_index():
    """torch.cdist's float32 arithmetic + torch.topk's own selection, both restated: the live reference's neighbour table,
    including the order inside groups of exactly equal float32 distances (a third of the rows of a ring scene)."""
    z = np.load(os.path.join(GOLDEN_DIR, "selection.npz"))
    for name, (flat, _, kn) in selection_cases().items():
        ref = z[f"{name}_nn"]
        if ref.shape[0] <= 1:
            continue
        idx = O.nearest_neighbors_cdist(flat, kn)
        D = np.sqrt(np.maximum(O.cdist_squared_f32(flat), np.float32(0.0)))
        np.fill_diagonal(D, np.inf)
        stable = np.argsort(D, axis=1, kind="stable")[:, :ref.shape[1]]
        tied = int((stable != ref).any(axis=1).sum())
        print(f"[knn] {name}: {int((idx != ref).any(axis=1).sum())} of {ref.shape[0]} rows differ from the reference "
              f"({tied} rows would with a lower-index-first tie rule)")
        assert np.array_equal(idx, ref), name


def test_topk_restatement_equals_torch_topk_on_tie_heavy_rows():
    """std::partial_sort (k * 64 <= n) and std::nth_element + std::sort (otherwise) as torch.topk's CPU kernel runs them,
    restated move for move: same indices in the same order as torch.topk on rows made of a few distinct values, with
    +inf and NaN entries."""
    import torch
    rs = np.random.RandomState(11)
    paths = {"partial_sort": 0, "nth_element": 0}
    for trial in range(600):
        n = int(rs.randint(2, 400)) if trial % 3 else int(rs.randint(300, 2500))
        k = int(rs.randint(1, min(n, 20)))
        row = rs.randint(0, int(rs.randint(1, 12)), size=n).astype(np.float32)
        if trial % 7 == 0:
            row[rs.randint(0, n)] = np.inf
        if trial % 11 == 0:
            row[rs.randint(0, n)] = np.nan
        want = torch.topk(torch.from_numpy(row)[None, :], k, largest=False, dim=1)[1][0].tolist()
        assert O.topk_smallest_like_torch(row, k) == want, (trial, n, k)
        paths["partial_sort" if k * 64 <= n else "nth_element"] += 1
    assert min(paths.values()) > 50, paths


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


def test_topk_restatement_equals_torch_topk_on_whole_matrices():
    """The reference calls torch.topk on the whole [n, n] distance matrix (rows run in parallel inside ATen): row by row the
    result must still be the restated selection, for lattice poses whose rows are full of exact ties, on both selection paths."""
    import torch
    rs = np.random.RandomState(12)
    for n, k in ((40, 8), (200, 4), (300, 4), (700, 10), (1200, 16)):
        flat = rs.randint(-2, 3, size=(n, 16)).astype(np.float32)
        flat[:, 12:] = [0, 0, 0, 1]
        m = torch.from_numpy(flat)
        dist = torch.cdist(m, m, p=2)
        dist.fill_diagonal_(float("inf"))
        want = torch.topk(dist, k, largest=False, dim=1)[1].numpy()
        D = dist.numpy()
        got = np.asarray([O.topk_smallest_like_torch(D[i], k) for i in range(n)])
        assert np.array_equal(got, want), (n, k, int((got != want).any(axis=1).sum()))
        # and the restated cdist arithmetic gives torch's distances bit for bit on these poses
        mine = np.sqrt(np.maximum(O.cdist_squared_f32(flat), np.float32(0.0)))
        np.fill_diagonal(mine, np.inf)
        assert np.array_equal(mine, D), (n, k)


def test_nearest_neighbours_small_scenes_equal_live_torch():
    """torch.cdist switches formulation at 25 rows (direct sum of squared differences below, matrix product above): the
    restatement follows, distances bit for bit where the direct kernel runs and tables index for index on both sides."""
    import torch
    from lichtfeld_densification_plugin_b200 import synth
    rs = np.random.RandomState(0)
    for n in (2, 3, 5, 12, 20, 25, 26, 30):
        sc = synth.make_scene(max(n, 3), "turbo", 0.5, 2)
        ring = np.stack([c.flat_pose() for c in sc.cameras], 0).astype(np.float32)[:n]
        rnd = rs.standard_normal((n, 16)).astype(np.float32)
        rnd[:, 12:] = [0, 0, 0, 1]
        for flat in (ring, rnd):
            m = torch.from_numpy(flat)
            d = torch.cdist(m, m, p=2)
            if n <= 25:
                assert np.array_equal(O.cdist_f32(flat), d.numpy()), n
            d.fill_diagonal_(float("inf"))
            k = max(1, min(4, n - 1))
            want = torch.topk(d, k, largest=False, dim=1)[1].numpy()
            assert np.array_equal(O.nearest_neighbors_cdist(flat, 4), want), n


def test_einsum_row_reduction_restated_bit_for_bit():
    """The first k-centres pick is argmax(np.einsum("nd,nd->n", Xn, Xn)): on symmetric rings the row norms tie and the pick is
    einsum's rounding.  The restated order (4 lanes, last vector first, multiply then add, (l0 + l1) + (l2 + l3)) against
    np.einsum itself; a mismatch here means numpy changed its einsum loops (the goldens would show it too)."""
    rs = np.random.RandomState(6)
    X = rs.standard_normal((20000, 16)).astype(np.float32)
    assert np.array_equal(O._einsum_row_sum16_f32(X * X), np.einsum("nd,nd->n", X, X))
    Y = (rs.randint(-3, 4, size=(5000, 16)) / np.float32(7.0)).astype(np.float32)      # many near-equal rows
    assert np.array_equal(O._einsum_row_sum16_f32(Y * Y), np.einsum("nd,nd->n", Y, Y))
