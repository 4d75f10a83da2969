"""dev tool: where the end-to-end (host buffers) step spends its time, and whether chunked pipelining helps."""
import sys, time
sys.path.insert(0, '/root/repo')
import torch
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
eng = DensifyEngine(dev)
cfg = PathConfig(matches_per_ref=WL["M"], seed=0)
cams = scene.cameras
def make_batch(cert_t, warp_t, image_t, lo=0, hi=R):
    b = eng.new_batch(H, W, wm, hm)
    for rp in range(lo, hi):
        ri, nb = tab[rp]
        b.add([cert_t[rp, k] for k in range(nn)], [warp_t[rp, k] for k in range(nn)], image_t[rp], cams[ri], [cams[j] for j in nb], rng_stream=rp)
    return b
sel_cap = eng.sel_capacity(cfg.matches_per_ref)
h_cert = cert.cpu().pin_memory(); h_img = image.cpu().pin_memory(); h_warp = warp.cpu().pin_memory()
d_cert = torch.empty_like(cert); d_img = torch.empty_like(image)

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e3

def h2d():
    d_cert.copy_(h_cert, non_blocking=True); d_img.copy_(h_img, non_blocking=True)
print("H2D cert+img ms", timeit(h2d), "GB/s", (h_cert.numel()*4 + h_img.numel()) / timeit(h2d) / 1e6)
b_dev = make_batch(d_cert, warp, d_img); descs_dev = eng.upload_descs(b_dev); out = eng.alloc_outputs(R, sel_cap)
print("densify all-device ms", timeit(lambda: eng.densify(b_dev, cfg, descs_dev=descs_dev, outputs=out)))
b_h = make_batch(d_cert, h_warp, d_img); descs_h = eng.upload_descs(b_h)
print("densify host-warp ms", timeit(lambda: eng.densify(b_h, cfg, descs_dev=descs_h, outputs=out)))
b_hh = make_batch(h_cert, h_warp, d_img); descs_hh = eng.upload_descs(b_hh)
def zc():
    d_img.copy_(h_img, non_blocking=True)
    eng.densify(b_hh, cfg, descs_dev=descs_hh, outputs=out)
print("img H2D + densify zero-copy cert + host-warp ms", timeit(zc))
cap = R * sel_cap
h_xyz = torch.empty((cap, 3)).pin_memory(); h_rgb = torch.empty((cap, 3)).pin_memory(); h_err = torch.empty((cap,)).pin_memory()
h_off = torch.empty((R + 1,), dtype=torch.int64).pin_memory()
def d2h():
    h_off.copy_(out.ref_offset, non_blocking=True); torch.cuda.current_stream().synchronize()
    n = int(h_off[-1])
    h_xyz[:n].copy_(out.xyz[:n], non_blocking=True); h_rgb[:n].copy_(out.rgb[:n], non_blocking=True); h_err[:n].copy_(out.err[:n], non_blocking=True)
    torch.cuda.current_stream().synchronize()
print("D2H ms", timeit(d2h))
def full():
    h2d(); eng.densify(b_h, cfg, descs_dev=descs_h, outputs=out); d2h()
print("e2e serial ms", timeit(full))
def full_zc():
    zc(); d2h()
print("e2e zero-copy-cert ms", timeit(full_zc))
# chunked pipeline: copy stream + compute stream
for nchunk in (2, 4, 8):
    bounds = [(R * i // nchunk, R * (i + 1) // nchunk) for i in range(nchunk)]
    batches = [make_batch(d_cert, h_warp, d_img, lo, hi) for lo, hi in bounds]
    descs = [eng.upload_descs(b) for b in batches]
    outs = [eng.alloc_outputs(hi - lo, sel_cap) for lo, hi in bounds]
    cs = torch.cuda.Stream(dev); evs = [torch.cuda.Event() for _ in bounds]
    h_offs = [torch.empty((hi - lo + 1,), dtype=torch.int64).pin_memory() for lo, hi in bounds]
    def piped():
        main = torch.cuda.current_stream()
        cs.wait_stream(main)
        with torch.cuda.stream(cs):
            for i, (lo, hi) in enumerate(bounds):
                d_cert[lo:hi].copy_(h_cert[lo:hi], non_blocking=True); d_img[lo:hi].copy_(h_img[lo:hi], non_blocking=True)
                evs[i].record(cs)
        for i in range(nchunk):
            main.wait_event(evs[i])
            eng.densify(batches[i], cfg, descs_dev=descs[i], outputs=outs[i])
            h_offs[i].copy_(outs[i].ref_offset, non_blocking=True)
        main.synchronize()
        base = 0
        for i in range(nchunk):
            n = int(h_offs[i][-1])
            h_xyz[base:base+n].copy_(outs[i].xyz[:n], non_blocking=True); h_rgb[base:base+n].copy_(outs[i].rgb[:n], non_blocking=True)
            h_err[base:base+n].copy_(outs[i].err[:n], non_blocking=True); base += n
        main.synchronize()
    print(f"e2e piped x{nchunk} ms", timeit(piped))
