"""cProfile of the drop-in entry point (core.pipeline.triangulate_refs) on the bench batch."""
import cProfile, pstats, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from lichtfeld_densification_plugin_b200 import synth
from lichtfeld_densification_plugin_b200.core import pipeline as PL
from lichtfeld_densification_plugin_b200.core.config import DensePipelineConfig
dev = torch.device("cuda", 0)
scene = synth.make_scene(185, "fast", 0.25, 4)
inputs = bench.SceneInputs(scene, list(range(scene.n_refs)), dev, seed=100)
cams = scene.cameras
cfg = DensePipelineConfig(output_path="/tmp/x.ply", matches_per_ref=10000)
ctx = PL._TriangulationContext(cameras=PL._build_camera_lookup(cams), config=cfg, matcher_sample_cap=0.9, w_match=scene.w_match, h_match=scene.h_match)
mrs = []
for i, (ri, nb) in enumerate(inputs.table):
    packed = PL._PackedReferenceBatch(ref_id=cams[ri].uid, ref_path="", imA_np=inputs.image[i], maskA_np=None, wA_cam=cams[ri].width,
                                      hA_cam=cams[ri].height, nn_ids=[cams[j].uid for j in nb], nn_masks=[None] * len(nb), nn_arrays=[])
    mrs.append(PL._MatchedReference(packed=packed, warp_list_cpu=[inputs.warp[i, k] for k in range(len(nb))],
                                    cert_list_cpu=[inputs.cert[i, k] for k in range(len(nb))], pair_index_by_nbr={}, image_by_nbr={}))
streams = [int(m.packed.ref_id) for m in mrs]
for _ in range(3):
    PL.triangulate_refs(mrs, ctx, rng_streams=streams)
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    pend = PL.submit_refs(mrs, ctx, rng_streams=streams)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
for _ in range(10):
    res = PL.collect_refs(pend)
t3 = time.perf_counter()
print(f"submit {1e2*(t1-t0):.3f} ms/call, collect {1e2*(t3-t2):.3f} ms/call")
pr = cProfile.Profile(); pr.enable()
for _ in range(10):
    PL.triangulate_refs(mrs, ctx, rng_streams=streams)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)
# the bench's own loop: per-call wall times (the result of the previous call is alive while the next one runs)
import time as _t
ts = []
for _ in range(12):
    torch.cuda.synchronize()
    t0 = _t.perf_counter()
    res = PL.triangulate_refs(mrs, ctx, rng_streams=streams)
    ts.append(round(1e3 * (_t.perf_counter() - t0), 2))
print("per-call ms:", ts)
