"""Dev tool: config-5 scene, 250 views cut into launches of PER views on the ring (3 in flight) against a sequential single-engine run."""
import sys, os
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import bench
from lichtfeld_densification_plugin_b200 import synth
from lichtfeld_densification_plugin_b200.engine import DensifyEngine, DensifyRing, PathConfig
dev = torch.device("cuda", 0)
if os.environ.get("WS_SLACK"):
    import ctypes as C
    from lichtfeld_densification_plugin_b200 import _native as N
    def _ensure(self, params):
        need = C.c_size_t(0)
        N.check(self.lib.ldp_workspace_bytes(C.byref(params), C.byref(need)), "ldp_workspace_bytes")
        k = int(os.environ["WS_SLACK"])
        if self._workspace is None or self._workspace.numel() < need.value * k:
            self._workspace = torch.zeros(int(need.value) * k, dtype=torch.uint8, device=self.device)
        return self._workspace
    DensifyEngine._ensure_workspace = _ensure
C5 = bench.CONFIG5
PER = int(sys.argv[1]); USE_RING = int(sys.argv[2]) if len(sys.argv) > 2 else 1
scene = synth.make_scene(C5["n_views"], C5["setting"], C5["ref_fraction"], C5["nn"])
R = scene.n_refs
inputs = bench.SceneInputs(scene, list(range(0, R)), dev, seed=500)
cfg = PathConfig(matches_per_ref=C5["M"], seed=5)
ring = DensifyRing(dev, 3)
eng0 = DensifyEngine(dev)
if len(sys.argv) > 3 and int(sys.argv[3]):        # prelude: what the bench did on these engines before (46 views at 512^2)
    sc2 = synth.make_scene(185, "fast", 0.25, 4)
    in2 = bench.SceneInputs(sc2, list(range(sc2.n_refs)), dev, seed=100)
    cfg2 = PathConfig(matches_per_ref=10000, seed=0)
    for e in ring.engines:
        b2 = in2.batch(e, sc2, 0, sc2.n_refs, stream_base=0)
        o2 = e.alloc_outputs(sc2.n_refs, e.sel_capacity(10000))
        for _ in range(2):
            e.prepare(b2, cfg2, outputs=o2).launch()
    torch.cuda.synchronize()
    print("prelude done")
sel_cap = eng0.sel_capacity(cfg.matches_per_ref)

def per_view(outs, chunks):
    res = {}
    for o, (a, b) in zip(outs, chunks):
        off = o.ref_offset.cpu().numpy(); ns = o.n_samples.cpu().numpy(); xyz = o.xyz.cpu().numpy(); st = o.status.cpu().numpy()
        uu = o.uniforms_used.cpu().numpy(); rr = o.rounds.cpu().numpy()
        for r in range(b - a):
            res[a + r] = (int(ns[r]), int(st[r]), xyz[off[r]:off[r + 1]].copy(), int(uu[r]), int(rr[r]))
    return res

# reference: sequential, one engine, 46 per launch
chunks0 = [(a, min(a + 46, R)) for a in range(0, R, 46)]
outs0 = []
for a, b in chunks0:
    batch = inputs.batch(eng0, scene, a, b, stream_base=0)
    o = eng0.alloc_outputs(b - a, sel_cap)
    eng0.prepare(batch, cfg, outputs=o).launch(); torch.cuda.synchronize()
    outs0.append(o)
ref = per_view(outs0, chunks0)

chunks = [(a, min(a + PER, R)) for a in range(0, R, PER)]
outs = [eng0.alloc_outputs(b - a, sel_cap) for a, b in chunks]
prepared = []
for c, (a, b) in enumerate(chunks):
    j = c % 3
    e = ring.engines[j] if USE_RING else eng0
    batch = inputs.batch(e, scene, a, b, stream_base=0)
    prepared.append((j, e.prepare(batch, cfg, descs_dev=e.upload_descs(batch), outputs=outs[c])))
NPASS = int(sys.argv[4]) if len(sys.argv) > 4 else 1
for trial in range(3):
    main = torch.cuda.current_stream(dev)
    for _ in range(NPASS):                 # passes back to back, no host synchronisation in between (what the bench does)
        if USE_RING:
            for st in ring.streams: st.wait_stream(main)
            for j, p in prepared:
                with torch.cuda.stream(ring.streams[j]): p.launch()
            for st in ring.streams: main.wait_stream(st)
        else:
            for j, p in prepared: p.launch()
    torch.cuda.synchronize()
    got = per_view(outs, chunks)
    bad = [v for v in range(R) if got[v][0] != ref[v][0] or got[v][1] != ref[v][1] or got[v][2].shape != ref[v][2].shape or not np.array_equal(got[v][2], ref[v][2])]
    print(f"trial {trial}: PER {PER} ring {USE_RING}: {len(bad)} views differ", bad[:12], [(v % PER, "n", got[v][0], ref[v][0], "kept", got[v][2].shape[0], ref[v][2].shape[0], "uniforms", got[v][3], ref[v][3], "rounds", got[v][4], ref[v][4]) for v in bad[:4]])
from lichtfeld_densification_plugin_b200.output import ConcatPlan, PackedCloud, concat_launches
cloud = PackedCloud(R * sel_cap, dev)
plan = ConcatPlan([o.xyz for o in outs], [o.rgb for o in outs], [o.err for o in outs], [o.ref_offset[o.n_refs:o.n_refs + 1] for o in outs], PER * sel_cap)
plan.run(cloud); torch.cuda.synchronize()
refc = concat_launches(outs0); torch.cuda.synchronize()
k, k0 = cloud.total_points(), refc.total_points()
print("concat: totals", k, k0, "xyz equal", bool(k == k0 and torch.equal(cloud.xyz[:k], refc.xyz[:k])), "rgb", bool(k == k0 and torch.equal(cloud.rgb[:k], refc.rgb[:k])), "err", bool(k == k0 and torch.equal(cloud.err[:k], refc.err[:k])))
if k == k0:
    d = (cloud.xyz[:k] != refc.xyz[:k]).any(dim=1).nonzero().flatten()
    print("rows differing:", d.numel(), d[:10].tolist(), "chunk row starts:", np.cumsum([0] + [int(o.ref_offset[o.n_refs].item()) for o in outs]).tolist())
