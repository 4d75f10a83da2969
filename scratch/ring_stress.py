"""Dev tool: launches in flight on the ring, several passes back to back, at several shapes; every launch must equal the same launch
run alone on one engine.  python scratch/ring_stress.py"""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from lichtfeld_densification_plugin_b200 import synth
from lichtfeld_densification_plugin_b200.engine import DensifyEngine, DensifyRing, PathConfig
dev = torch.device("cuda", 0)

def case(name, nviews, setting, frac, nn, cuts, M, no_filter=False, passes=4, reserve=16):
    scene = synth.make_scene(nviews, setting, frac, nn)
    R = scene.n_refs
    inputs = bench.SceneInputs(scene, list(range(R)), dev, seed=300)
    cfg = PathConfig(matches_per_ref=M, seed=3, no_filter=no_filter)
    eng0 = DensifyEngine(dev)
    sel_cap = eng0.sel_capacity(M)
    ref = []
    for lo, hi in cuts:
        o = eng0.alloc_outputs(hi - lo, sel_cap)
        eng0.prepare(inputs.batch(eng0, scene, lo, hi, stream_base=0), cfg, outputs=o).launch(); torch.cuda.synchronize()
        k = o.total_points()
        ref.append((o.n_samples.clone(), o.ref_offset.clone(), o.xyz[:k].clone(), o.rgb[:k].clone()))
    ring = DensifyRing(dev, 3, sm_reserve=reserve)
    outs = [eng0.alloc_outputs(hi - lo, sel_cap) for lo, hi in cuts]
    prepared = []
    for c, (lo, hi) in enumerate(cuts):
        e = ring.engines[c % 3]
        b = inputs.batch(e, scene, lo, hi, stream_base=0)
        prepared.append((c % 3, e.prepare(b, cfg, descs_dev=e.upload_descs(b), outputs=outs[c])))
    main = torch.cuda.current_stream(dev)
    bad_total = 0
    for trial in range(3):
        for _ in range(passes):
            for st in ring.streams: st.wait_stream(main)
            for j, p in prepared:
                with torch.cuda.stream(ring.streams[j]): p.launch()
            for st in ring.streams: main.wait_stream(st)
        torch.cuda.synchronize()
        for c, o in enumerate(outs):
            ns, off, xyz, rgb = ref[c]
            k = int(off[-1])
            ok = torch.equal(o.n_samples, ns) and torch.equal(o.ref_offset, off) and torch.equal(o.xyz[:k], xyz) and torch.equal(o.rgb[:k], rgb)
            bad_total += 0 if ok else 1
    print(f"{name}: R={R} cuts={len(cuts)} -> launches that differ from the sequential run: {bad_total} of {3 * len(cuts)}", flush=True)

case("headline 512^2, 46 views x 6 launches", 185, "fast", 0.25, 4, [(0, 46)] * 6, 10000)
case("headline 512^2, reserve 0", 185, "fast", 0.25, 4, [(0, 46)] * 6, 10000, reserve=0)
case("512^2, 23-view launches", 185, "fast", 0.25, 4, [(0, 23), (23, 46)] * 3, 10000)
case("512^2, 15/16-view launches", 185, "fast", 0.25, 4, [(0, 15), (15, 30), (31, 46)] * 2, 10000)
case("ROI no_filter 8 nn", 40, "fast", 0.8, 8, [(0, 16), (16, 32)] * 3, 10000, no_filter=True)
case("precise 1280^2, 12-view launches", 100, "precise", 0.25, 4, [(0, 12), (12, 24)] * 3, 10000)
case("base 640^2, 50-view launches", 1000, "base", 0.25, 4, [(a, a + 50) for a in range(0, 250, 50)] + [(0, 50)], 10000)
case("turbo small maps, M=2000", 60, "turbo", 0.5, 3, [(0, 30)] * 6, 2000)
