"""GPU: the output contract and the post-path reducers on the device produce the
reference's bytes / rows: against goldens frozen from the live reference writers, `_apply_point_cap` and
`_build_filtered_match_preview`, and against the oracle's struct.pack restatement on random data."""
import os

import numpy as np
import pytest
import torch

from oracle import densify_oracle as O
from tests.golden.make_output_golden import CAP_CASES, PREVIEW_CASES, cap_inputs, preview_inputs
from tests.helpers import GOLDEN_DIR

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def out_mod():
    from lichtfeld_densification_plugin_b200 import output
    return output


def _dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def test_records_equal_reference_writer_golden(out_mod, tmp_path):
    z = np.load(os.path.join(GOLDEN_DIR, "writers.npz"))
    xyz, rgb, err = _dev(z["xyz"]), _dev(z["rgb"]), _dev(z["err"])
    assert np.array_equal(out_mod.to_uint8_rgb(rgb).cpu().numpy(), z["rgb_u8"])
    out_mod.write_ply(str(tmp_path / "a.ply"), xyz, rgb)
    out_mod.write_points3D_bin(str(tmp_path / "a.bin"), xyz, rgb, err)
    out_mod.write_points3D_bin(str(tmp_path / "b.bin"), xyz, rgb, None)
    assert (tmp_path / "a.ply").read_bytes() == z["ply"].tobytes()
    assert (tmp_path / "a.bin").read_bytes() == z["bin"].tobytes()
    assert (tmp_path / "b.bin").read_bytes() == z["bin_noerr"].tobytes()


@pytest.mark.parametrize("n", [0, 1, 255, 256, 257, 1000, 100003])
def test_records_random_sizes_against_struct_pack(out_mod, n):
    rs = np.random.RandomState(n + 1)
    xyz = (rs.standard_normal((n, 3)) * 10).astype(np.float32)
    rgb = (rs.random_sample((n, 3)) * 1.2 - 0.1).astype(np.float32)          # also below 0 and above 1: clipped
    if n > 8:
        rgb[:8, 0] = (np.arange(8, dtype=np.float32) + 0.5) / 255.0          # .5 cases: half to even
    err = rs.random_sample((n,)).astype(np.float32)
    u8 = O.to_uint8_rgb(rgb)
    m = min(n, 3000)                                                         # the struct.pack oracle is slow: check a prefix ...
    ply = out_mod.ply_records(_dev(xyz), _dev(rgb)).cpu().numpy().tobytes()
    b3d = out_mod.points3d_records(_dev(xyz), _dev(rgb), _dev(err)).cpu().numpy().tobytes()
    assert len(ply) == 15 * n and len(b3d) == 43 * n
    ref_ply = O.ply_bytes(xyz[:m], u8[:m])
    head = len(ref_ply) - 15 * m
    assert ply[:15 * m] == ref_ply[head:]
    assert b3d[:43 * m] == O.points3d_bin_bytes(xyz[:m], u8[:m], err[:m])[8:]
    if n:                                                                    # ... and everything against the vectorised writer
        from lichtfeld_densification_plugin_b200.core import writers as W
        rec = np.frombuffer(ply, dtype=W._PLY_VERTEX)
        assert np.array_equal(rec["x"], xyz[:, 0]) and np.array_equal(rec["z"], xyz[:, 2]) and np.array_equal(rec["g"], u8[:, 1])
        rec = np.frombuffer(b3d, dtype=W._BIN_POINT)
        assert np.array_equal(rec["id"], np.arange(1, n + 1, dtype=np.uint64)) and np.array_equal(rec["err"], err.astype(np.float64))
        assert np.array_equal(rec["y"], xyz[:, 1].astype(np.float64)) and np.array_equal(rec["b"], u8[:, 2])


def test_records_device_count_unaligned_buffer_and_first_id(out_mod):
    n, k = 5000, 1777
    rs = np.random.RandomState(3)
    xyz, rgb = rs.standard_normal((n, 3)).astype(np.float32), rs.random_sample((n, 3)).astype(np.float32)
    want = out_mod.ply_records(_dev(xyz[:k]), _dev(rgb[:k])).cpu().numpy()
    buf = torch.full((n * 15 + 1,), 0xAB, dtype=torch.uint8, device="cuda")
    n_dev = torch.tensor([k], dtype=torch.int64, device="cuda")
    out_mod.ply_records(_dev(xyz), _dev(rgb), n_dev=n_dev, out=buf[1:])                # misaligned destination, device-side count
    got = buf.cpu().numpy()
    assert np.array_equal(got[1:1 + 15 * k], want) and np.all(got[1 + 15 * k:] == 0xAB) and got[0] == 0xAB
    a = out_mod.points3d_records(_dev(xyz[100:200]), _dev(rgb[100:200]), None, first_id=101).cpu().numpy()
    b = out_mod.points3d_records(_dev(xyz[:200]), _dev(rgb[:200]), None).cpu().numpy()
    assert np.array_equal(a, b[100 * 43:])                                              # a rank writing rows [100, 200)


def test_point_cap_equals_reference_golden(out_mod):
    z = np.load(os.path.join(GOLDEN_DIR, "output_reducers.npz"))
    for i, c in enumerate(CAP_CASES):
        xyz, rgb, err = cap_inputs(c["n"], c["data_seed"])
        a, b, e = out_mod.apply_point_cap(_dev(xyz), _dev(rgb), _dev(err), c["max_points"], c["seed"])
        assert np.array_equal(a.cpu().numpy(), z[f"cap{i}_xyz"]) and np.array_equal(b.cpu().numpy(), z[f"cap{i}_rgb"])
        assert np.array_equal(e.cpu().numpy(), z[f"cap{i}_err"])


def test_preview_subsample_equals_reference_golden(out_mod):
    z = np.load(os.path.join(GOLDEN_DIR, "output_reducers.npz"))
    for i, c in enumerate(PREVIEW_CASES):
        m, cn = preview_inputs(c["k"], c["data_seed"])
        a, b = out_mod.subsample_preview_matches(_dev(m), _dev(cn), c["ref_id"], c["nbr_id"], c["max_matches"])
        assert np.array_equal(a.cpu().numpy(), z[f"pv{i}_matches"]) and np.array_equal(b.cpu().numpy(), z[f"pv{i}_cert"])
    with pytest.raises(IndexError):
        out_mod.gather_rows(_dev(np.zeros((4, 2), np.float32)), torch.tensor([0, 4], device="cuda"))
    assert out_mod.gather_rows(_dev(np.arange(8, dtype=np.float32).reshape(4, 2)), torch.tensor([-1], device="cuda")).tolist() == [[6.0, 7.0]]


def test_incremental_ply_equals_rewriting_everything(out_mod, tmp_path):
    from lichtfeld_densification_plugin_b200.core import writers as W
    rs = np.random.RandomState(8)
    inc = out_mod.IncrementalPly()
    xs, cs = [], []
    for step, n in enumerate([300, 0, 1, 4097]):
        xyz, rgb = rs.standard_normal((n, 3)).astype(np.float32), rs.random_sample((n, 3)).astype(np.float32)
        xs.append(xyz); cs.append(rgb)
        inc.append(_dev(xyz), _dev(rgb))
        inc.emit(str(tmp_path / f"i{step}.ply"))
        W.write_ply(str(tmp_path / f"w{step}.ply"), np.concatenate(xs), W.to_uint8_rgb(np.concatenate(cs)))   # core/pipeline.py:523-526
        assert (tmp_path / f"i{step}.ply").read_bytes() == (tmp_path / f"w{step}.ply").read_bytes()


def test_path_outputs_to_ply_without_host_sync(out_mod):
    """The path's packed outputs go to PLY records with the device-side point count (ref_offset[-1])."""
    from lichtfeld_densification_plugin_b200 import synth
    from lichtfeld_densification_plugin_b200.engine import DensifyEngine, PathConfig
    from lichtfeld_densification_plugin_b200.core import writers as W
    eng = DensifyEngine()
    scene = synth.make_scene(12, "turbo", 0.25, 2)
    b = eng.new_batch(scene.H, scene.W, scene.w_match, scene.h_match)
    for rp in range(scene.n_refs):
        inp = synth.synth_ref_inputs(scene, rp, device=eng.device, cert_family="R", seed=2)
        k = len(inp["nbr_indices"])
        b.add([inp["cert"][q] for q in range(k)], [inp["warp"][q] for q in range(k)], inp["image"], scene.cameras[inp["ref_index"]],
              [scene.cameras[j] for j in inp["nbr_indices"]], rng_stream=rp)
    out = eng.densify(b, PathConfig(matches_per_ref=3000))
    cap = int(out.err.shape[0])
    rec = out_mod.ply_records(out.xyz, out.rgb, n=cap, n_dev=out.ref_offset[-1:])
    K = out.total_points()
    assert 0 < K < cap
    xyz, rgb = out.xyz[:K].cpu().numpy(), out.rgb[:K].cpu().numpy()
    want = np.empty(K, dtype=W._PLY_VERTEX)
    want["x"], want["y"], want["z"] = xyz[:, 0], xyz[:, 1], xyz[:, 2]
    u8 = W.to_uint8_rgb(rgb)
    want["r"], want["g"], want["b"] = u8[:, 0], u8[:, 1], u8[:, 2]
    assert rec[:15 * K].cpu().numpy().tobytes() == want.tobytes()


def test_run_dense_pipeline_live_updates_write_the_reference_files(tmp_path):
    """run_dense_pipeline (reference signature) with a synthetic match source: the PipelineResult equals the per-view
    calls, and every intermediate PLY (core/pipeline.py:508-532: one every viz_interval views) holds exactly the bytes
    write_ply(concat(xyz so far), to_uint8_rgb(concat(rgb so far))) would - from records packed once on the device."""
    from lichtfeld_densification_plugin_b200 import synth
    from lichtfeld_densification_plugin_b200.core import pipeline as P
    from lichtfeld_densification_plugin_b200.core import writers as W
    from lichtfeld_densification_plugin_b200.core.config import DensePipelineConfig
    scene = synth.make_scene(24, "turbo", ref_fraction=0.3, nn=3)
    cams = scene.cameras

    def match_source(rp):
        inp = synth.synth_ref_inputs(scene, rp, cert_family="R", seed=4)
        ri, nb = inp["ref_index"], inp["nbr_indices"]
        packed = P._PackedReferenceBatch(ref_id=cams[ri].uid, ref_path="", imA_np=inp["image"].numpy(), maskA_np=None,
                                         wA_cam=cams[ri].width, hA_cam=cams[ri].height, nn_ids=[cams[j].uid for j in nb],
                                         nn_masks=[None] * len(nb), nn_arrays=[None] * len(nb))
        return P._MatchedReference(packed=packed, warp_list_cpu=[inp["warp"][k] for k in range(len(nb))],
                                   cert_list_cpu=[inp["cert"][k] for k in range(len(nb))], pair_index_by_nbr={}, image_by_nbr={})
    refs = list(range(scene.n_refs))
    assert len(refs) >= 6
    results = {}
    for per_launch in (0, 4, 2):                      # one launch; 2 and 4 launches kept in flight on the ring of engines
        out_dir = tmp_path / f"run{per_launch}"
        cfg = DensePipelineConfig(output_path=str(out_dir / "dense.ply"), matches_per_ref=2000, viz_interval=2, refs_per_launch=per_launch,
                                  viz_every_emission=True)
        emitted = []
        res = P.run_dense_pipeline(cams, refs, None, cfg, on_sequential_viz=emitted.append, match_source=match_source,
                                   w_match=scene.w_match, h_match=scene.h_match)
        assert res.pairs_processed == len(refs) and res.xyz.shape[0] > 1000
        results[per_launch] = res
        assert [os.path.basename(p) for p in emitted] == [f"dense_intermediate_{k}.ply" for k in range(2, len(refs) + 1, 2)]
        # per-view sizes: the same views one at a time (Philox streams are keyed by the view, not by the launch)
        ctx = P._TriangulationContext(cameras=P._build_camera_lookup(cams), config=cfg, matcher_sample_cap=0.9,
                                      w_match=scene.w_match, h_match=scene.h_match)
        sizes = [P.triangulate_refs([match_source(r)], ctx, rng_streams=[r])[0].xyz.shape[0] for r in refs]
        assert sum(sizes) == res.xyz.shape[0]
        for k, path in zip(range(2, len(refs) + 1, 2), emitted):
            n = sum(sizes[:k])
            W.write_ply(str(out_dir / "want.ply"), res.xyz[:n], W.to_uint8_rgb(res.rgb[:n]))
            assert open(path, "rb").read() == (out_dir / "want.ply").read_bytes(), (per_launch, k)
    for per_launch in (4, 2):                         # cutting the views into launches changes nothing
        assert np.array_equal(results[0].xyz, results[per_launch].xyz) and np.array_equal(results[0].rgb, results[per_launch].rgb)
        assert np.array_equal(results[0].err, results[per_launch].err)
    # no callback / interval 0: nothing is written
    cfg = DensePipelineConfig(output_path=str(tmp_path / "none" / "dense.ply"), matches_per_ref=2000, viz_interval=0)
    P.run_dense_pipeline(cams, refs[:2], None, cfg, on_sequential_viz=emitted.append, match_source=match_source,
                         w_match=scene.w_match, h_match=scene.h_match)
    assert not (tmp_path / "none").exists()


@pytest.mark.parametrize("n,vs,scale255", [(1, 0.5, False), (2000, 0.25, False), (50000, 0.05, True), (30000, 100.0, False), (4097, 1e-3, False)])
def test_voxel_downsample_equals_the_restated_open3d_algorithm(out_mod, n, vs, scale255):
    """PARITY UNPINNED against Open3D itself (not installed); exact against the oracle's restatement of its algorithm:
    same voxels, f64 means accumulated in point order, first-appearance order."""
    rs = np.random.RandomState(n)
    xyz = (rs.standard_normal((n, 3)) * 2).astype(np.float32)
    rgb = rs.random_sample((n, 3)).astype(np.float32) * (255.0 if scale255 else 1.0)
    want_xyz, want_rgb = O.voxel_downsample(xyz, rgb, vs)
    got_xyz, got_rgb = out_mod.voxel_downsample(_dev(xyz), _dev(rgb), vs)
    assert got_xyz.shape == want_xyz.shape and 1 <= got_xyz.shape[0] <= n
    assert np.array_equal(got_xyz.cpu().numpy(), want_xyz) and np.array_equal(got_rgb.cpu().numpy(), want_rgb)
    if vs == 100.0:
        assert got_xyz.shape[0] == 1                      # everything in one voxel: 30 000 points summed in order
    # idempotence-like property: every output point lies inside the voxel of the points it averages
    again_xyz, _ = out_mod.voxel_downsample(got_xyz, got_rgb, vs)
    assert again_xyz.shape[0] <= got_xyz.shape[0]


def test_voxel_downsample_rejects_bad_sizes(out_mod):
    xyz = _dev(np.array([[0, 0, 0], [1e6, 0, 0]], np.float32))
    rgb = _dev(np.zeros((2, 3), np.float32))
    with pytest.raises(Exception):
        out_mod.voxel_downsample(xyz, rgb, 0.0)
    with pytest.raises(ValueError):
        out_mod.voxel_downsample(xyz, rgb, 1e-3)          # 1e9 voxels along x: does not fit 21 bits


def test_run_dense_pipeline_debug_state_and_latest_only_live_update(tmp_path):
    """run_dense_pipeline honours ``debug_state`` like the reference (core/pipeline.py:866-893): with it enabled, every launch
    collects the kept matches per neighbour and one MatchPreview per due pair is submitted - the same previews the
    per-view drop-in call yields; by default only the newest due intermediate PLY of each collected launch is written."""
    from lichtfeld_densification_plugin_b200 import synth
    from lichtfeld_densification_plugin_b200.core import pipeline as P
    from lichtfeld_densification_plugin_b200.core import writers as W
    from lichtfeld_densification_plugin_b200.core.config import DensePipelineConfig
    scene = synth.make_scene(24, "turbo", ref_fraction=0.3, nn=3)
    cams = scene.cameras
    counter = {"pairs": 0}
    made = {}

    def match_source(rp):
        if rp in made:
            return made[rp]
        inp = synth.synth_ref_inputs(scene, rp, cert_family="R", seed=4)
        ri, nb = inp["ref_index"], inp["nbr_indices"]
        nn_ids = [cams[j].uid for j in nb]
        packed = P._PackedReferenceBatch(ref_id=cams[ri].uid, ref_path=f"img/{ri}.png", imA_np=inp["image"].numpy(), maskA_np=None,
                                         wA_cam=cams[ri].width, hA_cam=cams[ri].height, nn_ids=nn_ids,
                                         nn_masks=[None] * len(nb), nn_arrays=[None] * len(nb))
        pair_idx = {}
        for u in nn_ids:
            counter["pairs"] += 1
            pair_idx[u] = counter["pairs"]
        made[rp] = P._MatchedReference(packed=packed, warp_list_cpu=[inp["warp"][k] for k in range(len(nb))],
                                       cert_list_cpu=[inp["cert"][k] for k in range(len(nb))], pair_index_by_nbr=pair_idx,
                                       image_by_nbr={u: np.zeros((4, 4, 3), np.uint8) + i for i, u in enumerate(nn_ids)})
        return made[rp]

    class DebugState:
        def __init__(self, enabled):
            self.enabled, self.previews, self.total, self.released = enabled, [], None, False
        def is_enabled(self): return self.enabled
        def is_auto_step(self): return True
        def set_total_pairs(self, n): self.total = n
        def submit_preview(self, p): self.previews.append(p)
        def release_waiters(self): self.released = True

    refs = list(range(scene.n_refs))
    nn_table = scene.nn_table
    cfg = DensePipelineConfig(output_path=str(tmp_path / "dbg" / "dense.ply"), matches_per_ref=2000, viz_interval=2, refs_per_launch=3,
                              nns_per_ref=3)
    ds = DebugState(True)
    emitted, progress = [], []
    res = P.run_dense_pipeline(cams, refs, nn_table, cfg, progress_callback=lambda p, m: progress.append(p),
                               on_sequential_viz=emitted.append, debug_state=ds, match_source=match_source,
                               w_match=scene.w_match, h_match=scene.h_match)
    assert ds.released and ds.total is not None and ds.total > 0
    assert len(progress) >= (len(refs) + 2) // 3                                       # one progress report per collected launch
    # the previews the per-view drop-in call gives for the same views (pairs 1, 4, 7, ... when auto-stepping)
    ctx = P._TriangulationContext(cameras=P._build_camera_lookup(cams), config=cfg, matcher_sample_cap=0.9,
                                  w_match=scene.w_match, h_match=scene.h_match)
    want = []
    for r in refs:
        mr = match_source(r)
        tri = P.triangulate_refs([mr], ctx, True, rng_streams=[r])[0]
        for nbr_id, m in tri.debug_matches_by_nbr.items():
            if mr.pair_index_by_nbr[nbr_id] % 3 == 1:
                want.append((mr.packed.ref_id, nbr_id, mr.pair_index_by_nbr[nbr_id], m, tri.debug_cert_by_nbr[nbr_id]))
    assert len(ds.previews) == len(want) > 0
    for pv, (rid, nid, pidx, m, c) in zip(ds.previews, want):
        assert (pv.ref_id, pv.nbr_id, pv.pair_index, pv.total_pairs) == (rid, nid, pidx, ds.total)
        assert pv.match_count == m.shape[0] and np.array_equal(pv.matches, m) and np.array_equal(pv.cert_norm, c)
        assert pv.ref_label == f"{rid}.png" or pv.ref_label.endswith(".png")
    # live update, default: the newest due file of every collected launch only, each byte-identical to the reference's rewrite
    names = [os.path.basename(p) for p in emitted]
    assert names == sorted(set(names), key=names.index) and 0 < len(names) <= (len(refs) + 2) // 3
    sizes = [P.triangulate_refs([match_source(r)], ctx, rng_streams=[r])[0].xyz.shape[0] for r in refs]
    for path in emitted:
        k = int(os.path.basename(path).split("_")[-1].split(".")[0])
        assert k % 2 == 0
        n = sum(sizes[:k])
        W.write_ply(str(tmp_path / "want.ply"), res.xyz[:n], W.to_uint8_rgb(res.rgb[:n]))
        assert open(path, "rb").read() == (tmp_path / "want.ply").read_bytes(), k
    # disabled debug state: nothing is collected, nothing submitted
    ds2 = DebugState(False)
    P.run_dense_pipeline(cams, refs[:3], nn_table, cfg, debug_state=ds2, match_source=match_source, w_match=scene.w_match, h_match=scene.h_match)
    assert not ds2.previews and ds2.released


def test_concat_points_equals_concatenation_without_host_round_trip(out_mod):
    """ldp_concat_points (reference core/pipeline.py:914-928, np.concatenate of the per-view arrays): segments with their own
    padded arrays and device-side counts land back to back, in segment order; offsets and total stay on the device."""
    rs = np.random.RandomState(3)
    dev = torch.device("cuda", 0)
    caps = [1000, 4096, 7, 2048, 1]
    counts = [1000, 37, 0, 2047, 1]
    xyz = [torch.from_numpy(rs.standard_normal((c, 3)).astype(np.float32)).to(dev) for c in caps]
    rgb = [torch.from_numpy(rs.random_sample((c, 3)).astype(np.float32)).to(dev) for c in caps]
    err = [torch.from_numpy(rs.random_sample((c,)).astype(np.float32)).to(dev) for c in caps]
    cnt = [torch.tensor([c], dtype=torch.int64, device=dev) for c in counts]
    plan = out_mod.ConcatPlan(xyz, rgb, err, cnt, max(caps))
    dst = out_mod.PackedCloud(sum(caps), dev)
    plan.run(dst)
    torch.cuda.synchronize()
    n = dst.total_points()
    assert n == sum(counts)
    assert plan.seg_offsets.cpu().tolist() == np.concatenate([[0], np.cumsum(counts)]).tolist()
    assert torch.equal(dst.xyz[:n], torch.cat([a[:c] for a, c in zip(xyz, counts)]))
    assert torch.equal(dst.rgb[:n], torch.cat([a[:c] for a, c in zip(rgb, counts)]))
    assert torch.equal(dst.err[:n], torch.cat([a[:c] for a, c in zip(err, counts)]))
    # a destination that is too small is filled to its capacity and never overrun
    small = out_mod.PackedCloud(1024, dev)
    guard = small.packed.clone()
    plan.run(small)
    torch.cuda.synchronize()
    assert small.total_points() == small.capacity
    assert torch.equal(small.xyz[:1000], xyz[0][:1000])


def test_concat_launches_equals_one_launch(out_mod):
    """Two launches over halves of a batch, concatenated on the device, give the cloud of the single launch (Philox streams
    are keyed by the view): what run over a cut into launches - or over ranks - reassembles."""
    from lichtfeld_densification_plugin_b200 import synth
    from lichtfeld_densification_plugin_b200.engine import DensifyEngine, PathConfig
    eng = DensifyEngine()
    scene = synth.make_scene(16, "turbo", ref_fraction=0.4, nn=3)
    cfg = PathConfig(matches_per_ref=3000, seed=11)
    inputs = [synth.synth_ref_inputs(scene, rp, device=eng.device, cert_family="R", seed=8) for rp in range(scene.n_refs)]

    def run(lo, hi):
        b = eng.new_batch(scene.H, scene.W, scene.w_match, scene.h_match)
        for rp in range(lo, hi):
            inp = inputs[rp]
            nn = len(inp["nbr_indices"])
            b.add([inp["cert"][k] for k in range(nn)], [inp["warp"][k] for k in range(nn)], inp["image"],
                  scene.cameras[inp["ref_index"]], [scene.cameras[j] for j in inp["nbr_indices"]], rng_stream=rp)
        return eng.densify(b, cfg, outputs=eng.alloc_outputs(hi - lo, eng.sel_capacity(cfg.matches_per_ref)))
    R = scene.n_refs
    whole = run(0, R)
    parts = [run(0, R // 2), run(R // 2, R)]
    cloud = out_mod.concat_launches(parts)
    torch.cuda.synchronize()
    n = whole.total_points()
    assert cloud.total_points() == n > 1000
    assert torch.equal(cloud.xyz[:n], whole.xyz[:n]) and torch.equal(cloud.rgb[:n], whole.rgb[:n]) and torch.equal(cloud.err[:n], whole.err[:n])
