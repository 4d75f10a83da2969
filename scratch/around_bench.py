"""dev tool: device time of the rows around the path next to the oracle port of the reference on the host."""
import sys, time
sys.path.insert(0, '/root/repo')
import numpy as np, torch
from lichtfeld_densification_plugin_b200 import output as OUT, synth
from lichtfeld_densification_plugin_b200.core import selection as SEL, writers as W
from oracle import densify_oracle as O

dev = torch.device('cuda', 0)
def gpu_ms(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
def cpu_ms(fn, n=1):
    t0 = time.perf_counter()
    for _ in range(n): fn()
    return (time.perf_counter() - t0) / n * 1e3

K = 416288                                   # kept points of one bench step
rs = np.random.RandomState(0)
xyz = (rs.standard_normal((K, 3)) * 3).astype(np.float32); rgb = rs.random_sample((K, 3)).astype(np.float32); err = rs.random_sample(K).astype(np.float32)
dx, dc, de = (torch.from_numpy(a).to(dev) for a in (xyz, rgb, err))
print(f"points: {K}")
t = gpu_ms(lambda: OUT.ply_records(dx, dc)); print(f"PLY records (15 B/pt)           device {t*1e3:8.1f} us  ({K*39/t/1e6:6.1f} GB/s in+out)")
t = gpu_ms(lambda: OUT.points3d_records(dx, dc, de)); print(f"points3D records (43 B/pt)      device {t*1e3:8.1f} us  ({K*71/t/1e6:6.1f} GB/s in+out)")
m = 20000
u8 = O.to_uint8_rgb(rgb)
tc = cpu_ms(lambda: O.ply_bytes(xyz[:m], u8[:m])) * K / m; print(f"reference write_ply loop (struct.pack per point, oracle port, extrapolated from {m}) host {tc:8.1f} ms")
tc = cpu_ms(lambda: (W.to_uint8_rgb(rgb), W.write_ply('/tmp/_a.ply', xyz, W.to_uint8_rgb(rgb)))); print(f"vectorised numpy writer (ours, host)   {tc:8.1f} ms")
t = gpu_ms(lambda: OUT.apply_point_cap(dx, dc, de, 100000, 0), n=5); tc = cpu_ms(lambda: O.apply_point_cap(xyz, rgb, err, 100000, 0), 3)
print(f"point cap to 100k: device path (host choice + H2D idx + gather) {t:7.2f} ms   reference numpy {tc:7.2f} ms")
sel = torch.from_numpy(np.random.default_rng(0).choice(K, 100000, replace=False)).to(dev)
t = gpu_ms(lambda: OUT.gather_rows(dx, sel)); print(f"  gather kernel alone (100k rows x 12 B) {t*1e3:8.1f} us")
for vs in (0.05, 0.5):
    t = gpu_ms(lambda: OUT.voxel_downsample(dx, dc, vs), n=5); tc = cpu_ms(lambda: O.voxel_downsample(xyz, rgb, vs))
    nv = OUT.voxel_downsample(dx, dc, vs)[0].shape[0]
    print(f"voxel filter v={vs}: {nv} voxels  device {t:7.3f} ms   numpy restatement {tc:8.1f} ms")
for nv, k in ((185, 46), (1000, 250)):
    scene = synth.make_scene(nv, "turbo", 0.25, 4)
    flat = np.stack([c.flat_pose() for c in scene.cameras], 0)
    fd = torch.from_numpy(flat.astype(np.float32)).to(dev)
    t = gpu_ms(lambda: SEL.select_cameras_kcenters_device(fd, k)); tc = cpu_ms(lambda: O.select_cameras_kcenters(flat, k), 2)
    import importlib
    print(f"k-centres {nv} views -> {k}: device {t*1e3:8.1f} us   numpy (explicit-order oracle) {tc:8.2f} ms")
    t = gpu_ms(lambda: SEL.nearest_neighbors_device(fd, 4)); tc = cpu_ms(lambda: O.nearest_neighbors_exact(flat, 4), 2)
    print(f"4 nearest neighbours of {nv} views: device {t*1e3:8.1f} us   numpy f64 {tc:8.2f} ms")
