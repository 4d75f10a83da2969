"""dev tool: resident throughput of the path at the BASELINE.json configs 2-5 shapes (one GPU)."""
import sys, time
sys.path.insert(0, '/root/repo')
import torch
from lichtfeld_densification_plugin_b200 import synth
from lichtfeld_densification_plugin_b200.engine import DensifyEngine, PathConfig
dev = torch.device('cuda', 0)
eng = DensifyEngine(dev)
CONFIGS = [("config2 fast 185v/46 refs x4", 185, "fast", 0.25, 4, False),
           ("config3 precise 1280^2 maps, 46 refs x4", 185, "precise", 0.25, 4, False),
           ("config4 roi 40v/32 refs x8 no_filter", 40, "fast", 0.8, 8, True),
           ("config5 base 640^2, 1000v/250 refs x4", 1000, "base", 0.25, 4, False)]
for name, nv, setting, frac, nn, nf in CONFIGS:
    scene = synth.make_scene(nv, setting, frac, nn)
    R = scene.n_refs
    b = eng.new_batch(scene.H, scene.W, scene.w_match, scene.h_match)
    keep = []
    for rp in range(R):
        inp = synth.synth_ref_inputs(scene, rp, device=dev, cert_family="R", seed=100)
        keep.append(inp)
        k = len(inp["nbr_indices"])
        b.add([inp["cert"][q] for q in range(k)], [inp["warp"][q] for q in range(k)], inp["image"], scene.cameras[inp["ref_index"]],
              [scene.cameras[j] for j in inp["nbr_indices"]], rng_stream=rp)
    cfg = PathConfig(matches_per_ref=10000, no_filter=nf)
    descs = eng.upload_descs(b)
    out = eng.alloc_outputs(R, eng.sel_capacity(10000))
    for _ in range(3): eng.densify(b, cfg, descs_dev=descs, outputs=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 30
    e0.record()
    for _ in range(n): eng.densify(b, cfg, descs_dev=descs, outputs=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    pts = out.total_points()
    inbytes = sum(x["cert"].numel() * 4 for x in keep)
    print(f"{name:44s} R={R:3d} {scene.H}x{scene.W}  {ms*1e3:8.1f} us/step  {pts/ms/1e6:7.2f} G pts/s  {R*nn/ms/1e3:6.2f} M pairs/s  cert read {inbytes/ms/1e6:6.0f} GB/s", flush=True)
    import ctypes as C
    eng.lib.ldp_profile_enable(1)
    eng.densify(b, cfg, descs_dev=descs, outputs=out)
    buf = (C.c_float * 64)(); nk = eng.lib.ldp_profile_read(buf, 64)
    print("      ", {eng.lib.ldp_profile_name(k).decode().replace("ldp_", "").replace("_kernel", ""): round(buf[k] * 1e3, 1) for k in range(nk)}, flush=True)
    eng.lib.ldp_profile_enable(0)
    del keep, b, out, descs
    torch.cuda.empty_cache()
