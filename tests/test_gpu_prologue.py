"""GPU: the certainty post-processing step that precedes the path (reference core/pipeline.py:405-430),
as a stage (ldp_postprocess_certainty) and fused into the path's first kernel (ldp_params.prologue)."""
import dataclasses

import numpy as np
import pytest
import torch

from tests import gpu_harness as G
from tests.helpers import golden_scene
from tests.test_oracle_prologue import PRO_CASES, load_prologue_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def engine():
    from lichtfeld_densification_plugin_b200.engine import DensifyEngine
    return DensifyEngine(torch.device("cuda", 0))


def _golden_inputs(name):
    c, z, mA, mBs = load_prologue_golden(name)
    c = dict(c, no_filter=False)
    scene = golden_scene(c)
    raw = dict(cert=torch.from_numpy(z["raw_cert"]), warp=torch.from_numpy(z["warp"]), image=torch.from_numpy(z["image"]),
               ref_index=int(z["ref_index"]), nbr_indices=[int(x) for x in z["nbr_indices"]], mask_a=mA, masks_b=mBs)
    post = dict(raw, cert=torch.from_numpy(z["cert_post"]), mask_a=None, masks_b=None)
    return c, z, scene, raw, post


def _same_run(a: G.GpuRun, b: G.GpuRun, r: int = 0):
    assert np.array_equal(a.sel_idx[r], b.sel_idx[r])
    assert np.array_equal(a.flags[r], b.flags[r])
    assert np.array_equal(a.xyzerr[r], b.xyzerr[r], equal_nan=True)
    assert np.array_equal(a.xyz[r], b.xyz[r]) and np.array_equal(a.rgb[r], b.rgb[r]) and np.array_equal(a.err[r], b.err[r])
    assert a.weight_sum[r] == b.weight_sum[r] and a.uniforms_used[r] == b.uniforms_used[r]
    if a.dbg_cert is not None:
        assert np.array_equal(a.dbg_cert[r], b.dbg_cert[r]) and np.array_equal(a.dbg_matches[r], b.dbg_matches[r])


@pytest.mark.parametrize("name", PRO_CASES)
def test_stage_matches_live_reference_golden(engine, name):
    """ldp_postprocess_certainty == the cert_list the unmodified reference hands to _triangulate_ref, bit for bit."""
    c, z, scene, raw, _ = _golden_inputs(name)
    dev = engine.device
    batch = engine.new_batch(scene.H, scene.W, scene.w_match, scene.h_match)
    cams = scene.cameras
    nn = len(raw["nbr_indices"])
    cert, warp = raw["cert"].to(dev), raw["warp"].to(dev)
    to_dev = lambda m: None if m is None else torch.from_numpy(np.ascontiguousarray(m)).to(dev)
    batch.add([cert[k] for k in range(nn)], [warp[k] for k in range(nn)], raw["image"].to(dev), cams[raw["ref_index"]],
              [cams[j] for j in raw["nbr_indices"]], mask_a=to_dev(raw["mask_a"]), masks_b=[to_dev(m) for m in raw["masks_b"]])
    got = engine.postprocess_certainty(batch, c["certainty_thresh"]).cpu().numpy()[0]
    assert np.array_equal(got[:nn], z["cert_post"], equal_nan=True)


@pytest.mark.parametrize("name", PRO_CASES)
def test_fused_prologue_equals_preprocessed_and_reference(engine, name):
    """Raw planes + masks through the fused kernels == the reference's processed planes through the plain path ==
    the live reference's _triangulate_ref output frozen in the golden file."""
    c, z, scene, raw, post = _golden_inputs(name)
    U = np.random.RandomState(int(z["mt_seed"])).random_sample(3 * c["M"])[None, :]
    s = [float(z["weight_sum"])]
    cfg = G.path_cfg(c)
    fused = G.run_gpu(engine, scene, [raw], dataclasses.replace(cfg, certainty_floor=c["certainty_thresh"]), uniforms=U,
                      weight_sums=s, collect_debug=True)
    plain = G.run_gpu(engine, scene, [post], cfg, uniforms=U, weight_sums=s, collect_debug=True)
    _same_run(fused, plain)
    assert np.array_equal(fused.sel_idx[0], z["sel_idx"])
    assert fused.xyz[0].shape == z["xyz"].shape
    np.testing.assert_allclose(fused.xyz[0], z["xyz"], rtol=G.XYZ_RTOL, atol=G.XYZ_ATOL)
    np.testing.assert_allclose(fused.rgb[0], z["rgb"], rtol=G.XYZ_RTOL, atol=G.XYZ_ATOL)
    np.testing.assert_allclose(fused.err[0], z["err"], rtol=1e-4, atol=G.ERR_ATOL)
    # and with s computed on the device (exact f64 sum) instead of the reference box's torch-CPU reduction
    fused2 = G.run_gpu(engine, scene, [raw], dataclasses.replace(cfg, certainty_floor=c["certainty_thresh"]), uniforms=U)
    plain2 = G.run_gpu(engine, scene, [post], cfg, uniforms=U)
    _same_run(fused2, plain2)


def _full_size_masked_inputs(scene, R, seed=7):
    from lichtfeld_densification_plugin_b200 import synth
    rs = np.random.RandomState(seed)
    hm, wm = scene.h_match, scene.w_match
    yy, xx = np.mgrid[0:hm, 0:wm]
    inputs = []
    for rp in range(R):
        inp = synth.synth_ref_inputs(scene, rp, cert_family="R", seed=60)
        inp["cert"] = inp["cert"] - 0.1 * torch.rand(inp["cert"].shape, generator=torch.Generator().manual_seed(rp))
        nn = len(inp["nbr_indices"])
        cx, cy, rad = rs.uniform(0.3, 0.7) * wm, rs.uniform(0.3, 0.7) * hm, rs.uniform(0.35, 0.5) * wm
        inp["mask_a"] = (((xx - cx) ** 2 + (yy - cy) ** 2) < rad ** 2).astype(np.uint8) if rp % 3 != 2 else None
        inp["masks_b"] = [(((xx + 3 * yy + 17 * k) % 97) > 9).astype(np.uint8) if (k + rp) % 2 == 0 else None for k in range(nn)]
        inputs.append(inp)
    return inputs


@pytest.mark.parametrize("setting", ["fast", "precise"])
def test_fused_prologue_full_size(engine, setting):
    """BASELINE shapes (512^2; 1280^2 maps with 800^2 masks): fused raw path == stage kernel + plain path, every output."""
    from lichtfeld_densification_plugin_b200 import synth
    scene = synth.make_scene(24, setting, ref_fraction=0.125, nn=4)
    R = 3 if setting == "fast" else 1
    inputs = _full_size_masked_inputs(scene, R)
    c = dict(M=10000, no_filter=False, wm=scene.w_match, hm=scene.h_match)
    cfg = G.path_cfg(c, seed=4)
    floor = 0.2
    fused = G.run_gpu(engine, scene, inputs, dataclasses.replace(cfg, certainty_floor=floor), rng_streams=list(range(R)),
                      collect_debug=True)
    # stage kernel -> processed planes -> plain path
    dev = engine.device
    cams = scene.cameras
    batch = engine.new_batch(scene.H, scene.W, scene.w_match, scene.h_match)
    to_dev = lambda m: None if m is None else torch.from_numpy(np.ascontiguousarray(m)).to(dev)
    for inp in inputs:
        nn = len(inp["nbr_indices"])
        cert, warp = inp["cert"].to(dev), inp["warp"].to(dev)
        batch.add([cert[k] for k in range(nn)], [warp[k] for k in range(nn)], inp["image"].to(dev), cams[inp["ref_index"]],
                  [cams[j] for j in inp["nbr_indices"]], mask_a=to_dev(inp["mask_a"]), masks_b=[to_dev(m) for m in inp["masks_b"]])
    post_planes = engine.postprocess_certainty(batch, floor).cpu()
    post_inputs = [dict(inp, cert=post_planes[i, :len(inp["nbr_indices"])], mask_a=None, masks_b=None) for i, inp in enumerate(inputs)]
    plain = G.run_gpu(engine, scene, post_inputs, cfg, rng_streams=list(range(R)), collect_debug=True)
    for r in range(R):
        _same_run(fused, plain, r)
        assert fused.xyz[r].shape[0] > 1000
    # the stage itself against the oracle's explicit restatement (first view)
    from oracle import densify_oracle as O
    inp = inputs[0]
    for k in range(len(inp["nbr_indices"])):
        want = O.certainty_prologue(inp["cert"][k].numpy(), inp["warp"][k].numpy(), inp["mask_a"], inp["masks_b"][k], floor)
        assert np.array_equal(post_planes[0, k].numpy(), want, equal_nan=True)


def test_collect_reference_matches_drop_in(engine):
    """core.pipeline._collect_reference_matches + _triangulate_ref (reference signatures, numpy global RNG stream) on raw
    matcher outputs reproduce what the live reference produced for the golden case."""
    from lichtfeld_densification_plugin_b200.core import pipeline as P
    from lichtfeld_densification_plugin_b200.core.config import DensePipelineConfig
    c, z, scene, raw, _ = _golden_inputs("masks_resized")
    cams = scene.cameras
    ri, nb = raw["ref_index"], raw["nbr_indices"]
    dev = engine.device

    class Matcher:                      # stands in for RomaMatcher: outputs stay on the GPU
        def match_grids_batch(self, imA, nn_images):
            return [(raw["warp"][k].to(dev), raw["cert"][k].to(dev)) for k in range(len(nb))]

    cfg = DensePipelineConfig(output_path="/tmp/x.ply", matches_per_ref=c["M"], certainty_thresh=c["certainty_thresh"],
                              rng_mode="numpy")
    packed = P._PackedReferenceBatch(ref_id=cams[ri].uid, ref_path="", imA_np=raw["image"].numpy(), maskA_np=raw["mask_a"],
                                     wA_cam=cams[ri].width, hA_cam=cams[ri].height, nn_ids=[cams[j].uid for j in nb],
                                     nn_masks=raw["masks_b"], nn_arrays=[np.zeros((c["hm"], c["wm"], 3), np.uint8)] * len(nb))
    mr, counter = P._collect_reference_matches(packed, Matcher(), cfg, 0, None)
    assert counter == len(nb) and mr.raw_certainty and mr.cert_list_cpu[0].is_cuda
    ctx = P._TriangulationContext(cameras=P._build_camera_lookup(cams), config=cfg, matcher_sample_cap=0.9,
                                  w_match=c["wm"], h_match=c["hm"])
    np.random.seed(int(z["mt_seed"]))
    tri = P._triangulate_ref(mr, ctx, collect_debug_matches=True)
    assert tri is not None and abs(tri.xyz.shape[0] - z["xyz"].shape[0]) <= 3      # s: exact f64 sum vs torch-CPU f32 sum
    if tri.xyz.shape == z["xyz"].shape:
        np.testing.assert_allclose(tri.xyz, z["xyz"], rtol=G.XYZ_RTOL, atol=G.XYZ_ATOL)
        np.testing.assert_allclose(tri.rgb, z["rgb"], rtol=G.XYZ_RTOL, atol=G.XYZ_ATOL)
        assert list(tri.debug_matches_by_nbr.keys()) == [int(u) for u in z["dbg_uids"]]


@pytest.mark.parametrize("case", range(8))
def test_randomised_raw_planes(engine, case):
    """Seeded random shapes with raw certainties, a random floor and random masks (also at a resolution different from
    the maps): the fused path against the oracle run on planes post-processed by the oracle's explicit restatement."""
    from lichtfeld_densification_plugin_b200 import synth
    from oracle import densify_oracle as O
    rs = np.random.RandomState(4000 + case)
    H = int(rs.choice([48, 64, 96, 120]))
    W = int(rs.choice([48, 60, 64, 128]))
    hm = int(rs.choice([H, max(16, (H * 5) // 8)]))
    wm = int(rs.choice([W, max(16, (W * 5) // 8)]))
    nn = int(rs.randint(1, 5))
    M = int(rs.choice([300, 1200, 2500]))
    floor = float(rs.choice([0.2, 0.35, 0.0]))
    scene = synth.make_scene(10, "turbo", ref_fraction=0.2, nn=nn)
    scene.H, scene.W, scene.h_match, scene.w_match = H, W, hm, wm
    c = dict(M=M, no_filter=False, wm=wm, hm=hm)
    inputs, ress = [], []
    U = np.stack([np.random.RandomState(case * 7 + r).random_sample(3 * M + 64) for r in range(scene.n_refs)])
    for rp in range(scene.n_refs):
        inp = synth.synth_ref_inputs(scene, rp, cert_family="T", seed=300 + case)
        inp["cert"] = inp["cert"] - 0.15 * torch.rand(inp["cert"].shape, generator=torch.Generator().manual_seed(case * 10 + rp))
        k = len(inp["nbr_indices"])
        inp["mask_a"] = (rs.rand(hm, wm) > 0.25).astype(np.uint8) if rs.rand() < 0.7 else None
        inp["masks_b"] = [(rs.rand(hm, wm) > 0.2).astype(np.uint8) if rs.rand() < 0.6 else None for _ in range(k)]
        post = torch.from_numpy(np.stack([O.certainty_prologue(inp["cert"][q].numpy(), inp["warp"][q].numpy(), inp["mask_a"],
                                                               inp["masks_b"][q], floor) for q in range(k)]))
        try:
            res = G.run_oracle_ref(scene, dict(inp, cert=post), c, uniforms=U[rp])
        except ValueError as exc:                      # masks left fewer positive weights than the draw size
            assert "Fewer non-zero" in str(exc)
            res = "fewer"
        inputs.append(inp)
        ress.append(res)
    g = G.run_gpu(engine, scene, inputs, dataclasses.replace(G.path_cfg(c), certainty_floor=floor), uniforms=U,
                  weight_sums=[res.taps["s"] if res not in (None, "fewer") else 0.0 for res in ress], collect_debug=True)
    for r, res in enumerate(ress):
        if res == "fewer":
            assert g.status[r] & 0xFF == 2 and g.xyz[r].shape[0] == 0            # LDP_REF_FEWER_NONZERO, view skipped
            continue
        if res is None:
            assert g.xyz[r].shape[0] == 0
            continue
        rep = G.compare_ref(g, r, res, c, scene)
        assert rep.ok(), (case, H, W, hm, wm, nn, M, floor, rep)
        assert g.uniforms_used[r] == res.taps["uniforms_used"] and g.rounds[r] == res.taps["rounds"]
