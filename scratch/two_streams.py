"""dev tool: consecutive steps (independent batches) alternated over S streams, each with its own engine/workspace/outputs."""
import sys, time
sys.path.insert(0, '/root/repo')
import torch
import bench
from lichtfeld_densification_plugin_b200 import synth
from lichtfeld_densification_plugin_b200.engine import DensifyEngine, PathConfig
dev = torch.device('cuda', 0)
WL = bench.WORKLOAD
scene = synth.make_scene(WL["n_views"], WL["setting"], WL["ref_fraction"], WL["nn"])
R, nn = scene.n_refs, scene.nn
cfg = PathConfig(matches_per_ref=WL["M"], seed=0)
inputs = [synth.synth_ref_inputs(scene, rp, device=dev, cert_family="R", seed=100) for rp in range(R)]
def mk():
    eng = DensifyEngine(dev)
    b = eng.new_batch(scene.H, scene.W, scene.w_match, scene.h_match)
    for rp, inp in enumerate(inputs):
        k = len(inp["nbr_indices"])
        b.add([inp["cert"][q] for q in range(k)], [inp["warp"][q] for q in range(k)], inp["image"], scene.cameras[inp["ref_index"]],
              [scene.cameras[j] for j in inp["nbr_indices"]], rng_stream=rp)
    return eng, b, eng.upload_descs(b), eng.alloc_outputs(R, eng.sel_capacity(cfg.matches_per_ref))
for S in (1, 2, 3):
    sets = [mk() for _ in range(S)]
    streams = [torch.cuda.Stream(dev) for _ in range(S)]
    def run(n):
        for i in range(n):
            eng, b, d, o = sets[i % S]
            with torch.cuda.stream(streams[i % S]):
                eng.densify(b, cfg, descs_dev=d, outputs=o)
    run(6); torch.cuda.synchronize()
    n = 300
    t0 = time.perf_counter(); run(n); torch.cuda.synchronize(); dt = (time.perf_counter() - t0) / n
    print(f"streams={S}: {dt*1e6:7.1f} us/step  points {sets[0][3].total_points()}")
    del sets
