"""Dev tool: results must not depend on what the workspace held before the call.  Launch, fill the workspace with garbage, launch again."""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from lichtfeld_densification_plugin_b200 import synth
from lichtfeld_densification_plugin_b200.engine import DensifyEngine, PathConfig
dev = torch.device("cuda", 0)
def run_case(nviews, setting, frac, nn, R_use, M, seed):
    scene = synth.make_scene(nviews, setting, frac, nn)
    inputs = bench.SceneInputs(scene, list(range(0, R_use)), dev, seed=500)
    cfg = PathConfig(matches_per_ref=M, seed=seed)
    eng = DensifyEngine(dev)
    sel_cap = eng.sel_capacity(M)
    batch = inputs.batch(eng, scene, 0, R_use, stream_base=0)
    def launch():
        out = eng.alloc_outputs(R_use, sel_cap)
        eng.prepare(batch, cfg, outputs=out).launch(); torch.cuda.synchronize()
        off = out.ref_offset.cpu().numpy()
        return out.n_samples.cpu().numpy().copy(), off.copy(), out.xyz.cpu().numpy()[:off[-1]].copy(), out.status.cpu().numpy().copy()
    ref = launch()
    for name, fill in (("0xFF", lambda w: w.fill_(255)), ("random", lambda w: w.copy_(torch.randint(0, 256, w.shape, dtype=torch.uint8, device=dev))),
                       ("0x7F", lambda w: w.fill_(127)), ("zeros", lambda w: w.zero_())):
        fill(eng._workspace); torch.cuda.synchronize()
        got = launch()
        same = all(np.array_equal(a, b) for a, b in zip(ref, got))
        bad = [r for r in range(R_use) if got[0][r] != ref[0][r] or (got[1][r + 1] - got[1][r]) != (ref[1][r + 1] - ref[1][r])]
        print(f"{setting} R={R_use} M={M}: workspace filled with {name}: identical {same}; views with other counts: {bad[:8]}", flush=True)
run_case(1000, "base", 0.25, 4, 42, 10000, 5)
run_case(185, "fast", 0.25, 4, 46, 10000, 0)
run_case(24, "precise", 0.125, 4, 3, 10000, 1)
