import sys; sys.path.insert(0, "/root/repo")
import numpy as np, torch
from lichtfeld_densification_plugin_b200 import synth
from lichtfeld_densification_plugin_b200.core import pipeline as P
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
cfg = DensePipelineConfig(output_path="/tmp/o/dense.ply", matches_per_ref=2000, viz_interval=2)
ctx = P._TriangulationContext(cameras=P._build_camera_lookup(cams), config=cfg, matcher_sample_cap=0.9, w_match=scene.w_match, h_match=scene.h_match)
for kw in (dict(), dict(ply_records=True)):
    errs = []
    outs = P.triangulate_refs([match_source(r) for r in range(3)], ctx, rng_streams=[0, 1, 2], errors=errs, **kw)
    print(kw, [None if o is None else o.xyz.shape for o in outs], errs)
