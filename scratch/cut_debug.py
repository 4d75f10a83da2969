"""Dev tool: the same views of the config-5 scene through launches of different sizes / SM reserves must give identical per-view results."""
import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from lichtfeld_densification_plugin_b200 import synth
from lichtfeld_densification_plugin_b200.engine import DensifyEngine, PathConfig
dev = torch.device("cuda", 0)
C5 = bench.CONFIG5
scene = synth.make_scene(C5["n_views"], C5["setting"], C5["ref_fraction"], C5["nn"])
NV = int(sys.argv[1]) if len(sys.argv) > 1 else 46
inputs = bench.SceneInputs(scene, list(range(0, NV)), dev, seed=500)
cfg = PathConfig(matches_per_ref=C5["M"], seed=5)
eng = DensifyEngine(dev)
sel_cap = eng.sel_capacity(cfg.matches_per_ref)

def run(lo, hi, reserve):
    eng.sm_reserve = reserve
    batch = inputs.batch(eng, scene, lo, hi, stream_base=0)
    out = eng.alloc_outputs(hi - lo, sel_cap)
    eng.prepare(batch, cfg, outputs=out).launch()
    torch.cuda.synchronize()
    off = out.ref_offset.cpu().numpy()
    ns = out.n_samples.cpu().numpy()
    xyz = out.xyz.cpu().numpy()
    return [(int(ns[r]), xyz[off[r]:off[r + 1]].copy()) for r in range(hi - lo)], eng.lib.ldp_debug_last_cluster()

base, cl = run(0, NV, 0)
print("base: views", NV, "reserve 0 cluster", cl)
for (lo, hi, rs) in [(0, NV, 16), (0, 42, 16), (0, 42, 0), (0, 32, 16), (0, 40, 16), (2, 42, 16), (0, 21, 16), (21, 42, 16)]:
    if hi > NV: continue
    got, cl = run(lo, hi, rs)
    bad = [lo + r for r in range(hi - lo) if got[r][0] != base[lo + r][0] or got[r][1].shape != base[lo + r][1].shape or not np.array_equal(got[r][1], base[lo + r][1])]
    print(f"views [{lo},{hi}) reserve {rs} cluster {cl}: {len(bad)} views differ {bad[:10]}", [(got[b - lo][0], base[b][0], got[b - lo][1].shape[0], base[b][1].shape[0]) for b in bad[:3]])
