"""Dev tool (build container only: needs /root/reference): random small configurations through the LIVE reference's
_triangulate_ref and through the oracle with the same np.random seed; outputs must be bit-identical."""
import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from oracle import ref_import, densify_oracle as O
from lichtfeld_densification_plugin_b200 import synth
from tests import gpu_harness as G
from tests.golden.make_golden import build_scene

ref = ref_import.import_reference(full_pipeline=True)
P = ref.pipeline
torch.set_num_threads(1)
rs = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 20
bad = 0
for t in range(N):
    H = int(rs.randint(24, 120)); W = int(rs.randint(24, 120))
    hm, wm = (H, W) if rs.rand() < 0.5 else (int(rs.randint(16, 100)), int(rs.randint(16, 100)))
    nn = int(rs.randint(1, 5)); M = int(rs.randint(1, min(3000, H * W // 3)))
    c = dict(H=H, W=W, hm=hm, wm=wm, nn=nn, M=M, fam="T" if rs.rand() < 0.5 else "R", no_filter=bool(rs.rand() < 0.2),
             seed=int(rs.randint(1, 10000)), sampson=float(rs.choice([5.0, 0.0, 1.0])), parallax=float(rs.choice([0.5, 0.0, 2.0])))
    scene = build_scene(c)
    inp = synth.synth_ref_inputs(scene, 0, cert_family=c["fam"], seed=c["seed"])
    cams = scene.cameras
    ri, nb = inp["ref_index"], inp["nbr_indices"]
    cfg = ref.config.DensePipelineConfig(output_path="/tmp/unused.ply", matches_per_ref=M, no_filter=c["no_filter"],
                                         sampson_thresh=c["sampson"], min_parallax_deg=c["parallax"])
    ctx = P._TriangulationContext(cameras=P._build_camera_lookup(cams), config=cfg, matcher_sample_cap=0.9, w_match=wm, h_match=hm)
    packed = P._PackedReferenceBatch(ref_id=cams[ri].uid, ref_path="", imA_np=inp["image"].numpy(), maskA_np=None,
                                     wA_cam=cams[ri].width, hA_cam=cams[ri].height, nn_ids=[cams[j].uid for j in nb],
                                     nn_masks=[None] * len(nb), nn_arrays=[None] * len(nb))
    mr = P._MatchedReference(packed=packed, warp_list_cpu=[inp["warp"][k] for k in range(len(nb))],
                             cert_list_cpu=[inp["cert"][k] for k in range(len(nb))], pair_index_by_nbr={}, image_by_nbr={})
    np.random.seed(c["seed"])
    try:
        out = P._triangulate_ref(mr, ctx, collect_debug_matches=False)
        err_ref = None
    except Exception as exc:
        out, err_ref = None, f"{type(exc).__name__}: {exc}"
    np.random.seed(c["seed"])
    try:
        res = G.run_oracle_ref(scene, inp, c)
        err_o = None
    except Exception as exc:
        res, err_o = None, f"{type(exc).__name__}: {exc}"
    if err_ref or err_o:
        same = (err_ref == err_o)
    elif out is None or res is None:
        same = (out is None) == (res is None or res.xyz.shape[0] == 0)
    else:
        same = (out.xyz.shape == res.xyz.shape and np.array_equal(out.xyz, res.xyz) and np.array_equal(out.rgb, res.rgb)
                and np.array_equal(out.err, res.err, equal_nan=True))
    if not same:
        bad += 1
        print("MISMATCH", t, c, "ref", None if out is None else out.xyz.shape, err_ref, "oracle", None if res is None else res.xyz.shape, err_o)
print(f"{N} random configurations, mismatches: {bad}")
