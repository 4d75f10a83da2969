"""CPU, build container only: the oracle restatement vs the LIVE unmodified reference, bit for bit."""
import numpy as np
import pytest
import torch

from oracle import densify_oracle as O
from oracle import ref_import
from lichtfeld_densification_plugin_b200 import synth
from tests.helpers import oracle_cam

pytestmark = pytest.mark.skipif(not ref_import.reference_available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def ref():
    return ref_import.import_reference(full_pipeline=True)


@pytest.mark.parametrize("fam,no_filter,nn", [("T", False, 3), ("R", False, 2), ("T", True, 3), ("R", True, 1)])
def test_triangulate_ref_bit_identical(ref, fam, no_filter, nn):
    P = ref.pipeline
    scene = synth.make_scene(10, "turbo", ref_fraction=0.2, nn=nn)
    scene.H = scene.W = 128
    scene.h_match = scene.w_match = 96
    cams = scene.cameras
    for rp in range(scene.n_refs):
        inp = synth.synth_ref_inputs(scene, rp, cert_family=fam, seed=3)
        ri, nb = inp["ref_index"], inp["nbr_indices"]
        cfg = ref.config.DensePipelineConfig(output_path="/tmp/x.ply", no_filter=no_filter, matches_per_ref=5000)
        ctx = P._TriangulationContext(cameras=P._build_camera_lookup(cams), config=cfg, matcher_sample_cap=0.9,
                                      w_match=scene.w_match, h_match=scene.h_match)
        packed = P._PackedReferenceBatch(ref_id=cams[ri].uid, ref_path="", imA_np=inp["image"].numpy(), maskA_np=None,
                                         wA_cam=cams[ri].width, hA_cam=cams[ri].height,
                                         nn_ids=[cams[j].uid for j in nb], nn_masks=[None] * len(nb),
                                         nn_arrays=[None] * len(nb))
        mr = P._MatchedReference(packed=packed, warp_list_cpu=[inp["warp"][k] for k in range(len(nb))],
                                 cert_list_cpu=[inp["cert"][k] for k in range(len(nb))],
                                 pair_index_by_nbr={}, image_by_nbr={})
        np.random.seed(40 + rp)
        want = P._triangulate_ref(mr, ctx, collect_debug_matches=True)
        ocfg = O.OracleConfig(matches_per_ref=5000, no_filter=no_filter, w_match=scene.w_match, h_match=scene.h_match)
        got = O.triangulate_ref([inp["cert"][k] for k in range(len(nb))], [inp["warp"][k] for k in range(len(nb))],
                                inp["image"].numpy(), oracle_cam(cams[ri]), [oracle_cam(cams[j]) for j in nb], ocfg,
                                rng=np.random.RandomState(40 + rp), collect_debug=True)
        assert np.array_equal(want.xyz, got.xyz)
        assert np.array_equal(want.rgb, got.rgb)
        assert np.array_equal(want.err, got.err)
        assert list(want.debug_matches_by_nbr.keys()) == list(got.debug_matches_by_nbr.keys())
        for uid in want.debug_matches_by_nbr:
            assert np.array_equal(want.debug_matches_by_nbr[uid], got.debug_matches_by_nbr[uid])
            assert np.array_equal(want.debug_cert_by_nbr[uid], got.debug_cert_by_nbr[uid])


def test_sampler_and_geometry_functions(ref):
    rs = np.random.RandomState(0)
    cert = torch.from_numpy(rs.random_sample((90, 120)).astype(np.float32))
    np.random.seed(5)
    want = ref.sampling.select_samples_with_coverage(cert, 2000)
    got = O.select_samples(cert, 2000, rng=np.random.RandomState(5))
    assert np.array_equal(want, got)
    cams = synth.make_orbit_cameras(5)
    a, b = cams[0], cams[1]
    F1 = ref.geometry.fundamental_from_world2cam(a.K, a.R, a.t, b.K, b.R, b.t)
    F2 = O.fundamental_matrix(a.K, a.R, a.t, b.K, b.R, b.t)
    assert np.array_equal(F1, F2) and F1.dtype == F2.dtype
    uv1 = (rs.random_sample((500, 2)) * 800).astype(np.float32)
    uv2 = (uv1 + rs.standard_normal((500, 2)) * 5).astype(np.float32)
    assert np.array_equal(ref.geometry.sampson_error(F1, uv1, uv2), O.sampson_distance(F2, uv1, uv2))
    X1 = ref.geometry.dlt_triangulate_batch(a.P, b.P, uv1, uv2)
    X2 = O.dlt_points(a.P, b.P, uv1, uv2)
    assert np.array_equal(X1, X2)
    assert np.array_equal(ref.geometry.dlt_triangulate_batch(a.P, b.P, uv1[:1], uv2[:1]), O.dlt_points(a.P, b.P, uv1[:1], uv2[:1]))
    assert np.array_equal(ref.geometry.reprojection_errors(a.P, X1, uv1), O.reprojection_error(a.P, X2, uv1))
    assert np.array_equal(ref.geometry.cheirality_mask(b.P, X1), O.in_front(b.P, X2))
    assert np.array_equal(ref.geometry.parallax_mask(a.C, b.C, X1, 0.5), O.parallax_ok(a.C, b.C, X2, 0.5))
