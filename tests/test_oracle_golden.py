"""CPU: the oracle restatement replays the golden vectors made from the live reference."""
import numpy as np
import pytest

from oracle import densify_oracle as O
from tests.helpers import GOLDEN_CASES, GOLDEN_DIR, load_golden, run_oracle


def test_golden_cases_exist():
    assert len(GOLDEN_CASES) >= 6


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_replays_golden(name):
    c, scene, inp, z = load_golden(name)
    res = run_oracle(c, scene, inp, rng=np.random.RandomState(int(z["mt_seed"])), collect_debug=True,
                     s_override=z["weight_sum"])
    assert res is not None
    if not c["no_filter"]:
        # sorted sampler output is exact given the recorded f32 weight sum
        assert np.array_equal(res.sel_idx, z["sel_idx"])
    else:
        # top-M by capped certainty: family T has no ties, so even the unstable argsort is pinned
        assert np.array_equal(res.sel_idx, z["sel_idx"])
    assert res.xyz.shape == z["xyz"].shape
    # same box => bit-identical; another CPU may differ in LAPACK/BLAS last bits
    np.testing.assert_allclose(res.xyz, z["xyz"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(res.rgb, z["rgb"], rtol=0, atol=1e-7)
    np.testing.assert_allclose(res.err, z["err"], rtol=1e-3, atol=2e-4)
    assert sorted(res.debug_matches_by_nbr.keys()) == sorted(int(u) for u in z["dbg_uids"])
    for uid in res.debug_matches_by_nbr:
        np.testing.assert_allclose(res.debug_matches_by_nbr[uid], z[f"dbg_matches_{uid}"], rtol=0, atol=1e-4)
        np.testing.assert_allclose(res.debug_cert_by_nbr[uid], z[f"dbg_cert_{uid}"], rtol=0, atol=1e-7)


@pytest.mark.parametrize("name", [n for n in GOLDEN_CASES if "nofilter" not in n])
def test_explicit_uniform_stream_equals_global_rng(name):
    """The restated legacy choice on an explicit MT19937 stream == np.random.choice."""
    c, scene, inp, z = load_golden(name)
    U = np.random.RandomState(int(z["mt_seed"])).random_sample(3 * c["M"])
    res = run_oracle(c, scene, inp, uniforms=U, s_override=z["weight_sum"], keep_taps=True)
    assert np.array_equal(res.sel_idx, z["sel_idx"])
    assert res.taps["rounds"] >= 1 and res.taps["uniforms_used"] >= int(c["M"] * 0.85)


@pytest.mark.parametrize("n,size,seed", [(5000, 1200, 0), (40000, 8500, 1), (1000, 1000, 2), (64, 0, 3)])
def test_legacy_choice_restatement_vs_numpy(n, size, seed):
    rs = np.random.RandomState(100 + seed)
    w = rs.random_sample(n).astype(np.float32) + np.float32(0.2)
    w[rs.random_sample(n) < 0.1] = 0
    if size == n:
        w = np.abs(w) + np.float32(0.1)
    p = (w / w.sum()).astype(np.float32)
    want = np.random.RandomState(seed).choice(n, size=size, replace=False, p=p)
    U = np.random.RandomState(seed).random_sample(4 * max(size, 1))
    got, used, rounds = O.legacy_choice_no_replace(p, size, U)
    assert np.array_equal(got, want)
    # stream position after the call matches numpy's
    a = np.random.RandomState(seed)
    a.choice(n, size=size, replace=False, p=p)
    assert a.random_sample() == U[used]


def test_legacy_choice_errors():
    p = np.zeros(100, dtype=np.float32)
    p[:10] = 0.1
    with pytest.raises(ValueError, match="Fewer non-zero"):
        O.legacy_choice_no_replace(p, 20, np.zeros(100))
    with pytest.raises(ValueError, match="larger sample"):
        O.legacy_choice_no_replace(p, 200, np.zeros(1000))


def test_writers_golden():
    z = np.load(f"{GOLDEN_DIR}/writers.npz")
    u8 = O.to_uint8_rgb(z["rgb"])
    assert np.array_equal(u8, z["rgb_u8"])
    assert O.ply_bytes(z["xyz"], u8) == z["ply"].tobytes()
    assert O.points3d_bin_bytes(z["xyz"], u8, z["err"]) == z["bin"].tobytes()
    assert O.points3d_bin_bytes(z["xyz"], u8, None) == z["bin_noerr"].tobytes()
