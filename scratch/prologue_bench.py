"""dev tool: cost of the fused certainty post-processing at the bench workload (46 views, 512^2, 4 neighbours)."""
import sys, dataclasses
sys.path.insert(0, '/root/repo')
import numpy as np, torch, ctypes as C
import bench
from lichtfeld_densification_plugin_b200 import synth
from lichtfeld_densification_plugin_b200.engine import DensifyEngine, PathConfig
dev = torch.device('cuda', 0)
WL = bench.WORKLOAD
scene = synth.make_scene(WL["n_views"], WL["setting"], WL["ref_fraction"], WL["nn"])
R, nn, H, W = scene.n_refs, scene.nn, scene.H, scene.W
hm, wm = scene.h_match, scene.w_match
cert = torch.empty((R, nn, H, W), dtype=torch.float32, device=dev)
warp = torch.empty((R, nn, H, W, 4), dtype=torch.float32, device=dev)
image = torch.empty((R, hm, wm, 3), dtype=torch.uint8, device=dev)
tab = []
for rp in range(R):
    inp = synth.synth_ref_inputs(scene, rp, device=dev, cert_family="R", seed=100)
    cert[rp], warp[rp], image[rp] = inp["cert"], inp["warp"], inp["image"]
    tab.append((inp["ref_index"], inp["nbr_indices"]))
yy, xx = np.mgrid[0:hm, 0:wm]
mA = torch.from_numpy((((xx - 0.5 * wm) ** 2 + (yy - 0.5 * hm) ** 2) < (0.45 * wm) ** 2).astype(np.uint8)).to(dev)
mB = torch.from_numpy((((xx + 3 * yy) % 97) > 9).astype(np.uint8)).to(dev)
eng = DensifyEngine(dev)
cams = scene.cameras
def make_batch(use_a, use_b):
    b = eng.new_batch(H, W, wm, hm)
    for rp in range(R):
        ri, nb = tab[rp]
        b.add([cert[rp, k] for k in range(nn)], [warp[rp, k] for k in range(nn)], image[rp], cams[ri], [cams[j] for j in nb],
              rng_stream=rp, mask_a=mA if use_a else None, masks_b=[mB if use_b else None] * nn)
    return b
sel_cap = eng.sel_capacity(WL["M"])
out = eng.alloc_outputs(R, sel_cap)
for name, floor, ua, ub in (("plain", None, 0, 0), ("floor", 0.2, 0, 0), ("floor+maskA", 0.2, 1, 0), ("floor+maskA+maskB", 0.2, 1, 1)):
    cfg = PathConfig(matches_per_ref=WL["M"], seed=0, certainty_floor=floor)
    b = make_batch(ua, ub); descs = eng.upload_descs(b)
    for _ in range(5): eng.densify(b, cfg, descs_dev=descs, outputs=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): eng.densify(b, cfg, descs_dev=descs, outputs=out)
    e1.record(); torch.cuda.synchronize()
    eng.lib.ldp_profile_enable(1)
    eng.densify(b, cfg, descs_dev=descs, outputs=out)
    buf = (C.c_float * 64)(); n = eng.lib.ldp_profile_read(buf, 64)
    eng.lib.ldp_profile_enable(0)
    print(f"{name:20s} step {e0.elapsed_time(e1)/50*1e3:7.1f} us  stream {buf[0]*1e3:6.1f} us  pts {out.total_points()}")
