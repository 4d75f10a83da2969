"""dev tool: a few steps of the config-4 shape (32 views x 8 neighbours, 512^2, no_filter) for ncu.  argv[1] = cert family (R|T)."""
import sys
sys.path.insert(0, '/root/repo')
import torch
from lichtfeld_densification_plugin_b200 import synth
from lichtfeld_densification_plugin_b200.engine import DensifyEngine, PathConfig
fam = sys.argv[1] if len(sys.argv) > 1 else "R"
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device('cuda', 0)
eng = DensifyEngine(dev)
scene = synth.make_scene(40, "fast", 0.8, 8)
b = eng.new_batch(scene.H, scene.W, scene.w_match, scene.h_match); keep = []
for rp in range(scene.n_refs):
    inp = synth.synth_ref_inputs(scene, rp, device=dev, cert_family=fam, seed=100); keep.append(inp); k = len(inp["nbr_indices"])
    b.add([inp["cert"][q] for q in range(k)], [inp["warp"][q] for q in range(k)], inp["image"], scene.cameras[inp["ref_index"]],
          [scene.cameras[j] for j in inp["nbr_indices"]], rng_stream=rp)
cfg = PathConfig(matches_per_ref=10000, no_filter=True)
descs = eng.upload_descs(b); out = eng.alloc_outputs(scene.n_refs, eng.sel_capacity(10000))
for _ in range(steps): eng.densify(b, cfg, descs_dev=descs, outputs=out)
torch.cuda.synchronize()
print("points", out.total_points())
