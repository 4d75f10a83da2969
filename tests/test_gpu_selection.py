"""GPU: pair generation on the device against the live reference's frozen outputs and the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import densify_oracle as O
from tests.golden.make_selection_golden import selection_cases
from tests.helpers import GOLDEN_DIR
from tests.test_oracle_selection import nn_equivalent

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def S():
    from lichtfeld_densification_plugin_b200.core import selection
    return selection


def test_kcenters_equal_reference_golden(S):
    z = np.load(os.path.join(GOLDEN_DIR, "selection.npz"))
    for name, (flat, kc, _) in selection_cases().items():
        got = S.select_cameras_kcenters(flat, kc)
        assert isinstance(got, list) and got == z[f"{name}_centers"].tolist(), name
        _, order_dev = S.select_cameras_kcenters_device(flat, kc)
        assert order_dev.cpu().tolist() == O.select_cameras_kcenters(flat, kc)[1], name           # same greedy sequence


def test_kcenters_every_k_and_more_views_than_threads(S):
    rs = np.random.RandomState(5)
    flat = rs.standard_normal((2500, 16)).astype(np.float32)                                    # 3 views per thread
    flat[:, 12:] = [0, 0, 0, 1]
    for k in (1, 2, 17, 2500, 9999):
        want, order = O.select_cameras_kcenters(flat, k)
        assert S.select_cameras_kcenters(flat, k) == want, k
    big = np.concatenate([flat, flat[::-1] * np.float32(1.5)])[:4000]                          # beyond the shared-memory copy: global rows
    for k in (3, 40):
        assert S.select_cameras_kcenters(big, k) == O.select_cameras_kcenters(big, k)[0], k
    small = flat[:5]
    for k in range(1, 7):
        assert S.select_cameras_kcenters(small, k) == O.select_cameras_kcenters(small, k)[0]


def test_nearest_neighbours_equal_reference_golden(S):
    """Index-identical to the live reference's frozen tables: the kernel mirrors torch.cdist's float32 arithmetic AND runs
    torch.topk's own selection (libstdc++ partial_sort / nth_element restated), so even exactly tied distances - the left /
    right neighbours of ring cameras, a third of the rows there - come out in the reference's order."""
    z = np.load(os.path.join(GOLDEN_DIR, "selection.npz"))
    for name, (flat, _, kn) in selection_cases().items():
        got = S.nearest_neighbors(flat, kn)
        ref = z[f"{name}_nn"]
        assert got.dtype == np.int64 and got.shape == ref.shape, name
        print(f"[knn] {name}: {int((got != ref).any(axis=1).sum())} of {ref.shape[0]} rows differ from the reference")
        assert np.array_equal(got, ref), name
        assert np.array_equal(got, O.nearest_neighbors_cdist(flat, kn)), name


def test_nearest_neighbours_tie_heavy_rows_both_selection_paths(S):
    """Poses on a small integer lattice: most distances of a row are exactly tied.  Every k from 1 to 16 at sizes on both
    sides of k * 64 <= n (streamed heap select vs selection in shared memory) and beyond one staging chunk."""
    rs = np.random.RandomState(3)
    for n in (2, 3, 17, 64, 65, 255, 256, 257, 700, 1023, 1024, 1025, 2049, 5000):
        flat = rs.randint(-2, 3, size=(n, 16)).astype(np.float32)
        flat[:, 12:] = [0, 0, 0, 1]
        for k in sorted({1, 2, 3, 4, 7, 15, 16} & set(range(1, n))):
            want = O.nearest_neighbors_cdist(flat, k)
            got = S.nearest_neighbors(flat, k)
            assert np.array_equal(got, want), (n, k, int((got != want).any(axis=1).sum()))


def test_pairs_feed_the_path(S):
    """Selection -> neighbour table -> the views a launch processes (the device-resident front of the 1000-view config)."""
    from lichtfeld_densification_plugin_b200 import synth
    scene = synth.make_scene(40, "turbo", 0.25, 3)
    flat = np.stack([c.flat_pose() for c in scene.cameras], 0)
    refs = S.select_cameras_kcenters(flat, 10)
    nn = S.nearest_neighbors(flat, 3)
    assert len(refs) == 10 and nn.shape == (40, 3)
    assert S._estimate_total_pairs(refs, nn, list(range(40)), 3) == 30
    with pytest.raises(Exception):
        S.nearest_neighbors(flat, 17) if False else S.nearest_neighbors_device(np.zeros((40, 15), np.float32), 3)


def test_nearest_neighbours_small_scenes_take_torchs_direct_cdist(S):
    """Up to 25 views torch.cdist runs its direct kernel, not the matrix product: the table must follow (12-view ring: 4 rows
    differ between the two formulations).  Rings and random poses on both sides of the switch."""
    from lichtfeld_densification_plugin_b200 import synth
    rs = np.random.RandomState(4)
    for n in (2, 3, 5, 12, 20, 25, 26, 30):
        sc = synth.make_scene(max(n, 3), "turbo", 0.5, 2)
        ring = np.stack([c.flat_pose() for c in sc.cameras], 0).astype(np.float32)[:n]
        rnd = rs.standard_normal((n, 16)).astype(np.float32)
        rnd[:, 12:] = [0, 0, 0, 1]
        for flat in (ring, rnd):
            for k in (1, 4):
                assert np.array_equal(S.nearest_neighbors(flat, k), O.nearest_neighbors_cdist(flat, k)), (n, k)
