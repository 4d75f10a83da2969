"""Phase clocks of the front kernel (build with -DLDP_FRONT_CLOCKS): sums over all CTAs of thread 0's cycles per phase."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lichtfeld_densification_plugin_b200 import synth
from lichtfeld_densification_plugin_b200.engine import DensifyEngine, PathConfig
dev = torch.device("cuda", 0)
scene = synth.make_scene(185, "fast", 0.25, 4)
R = scene.n_refs
eng = DensifyEngine(dev)
cfg = PathConfig(matches_per_ref=10000, seed=0)
cams = scene.cameras
b = eng.new_batch(scene.H, scene.W, scene.w_match, scene.h_match)
for rp in range(R):
    inp = synth.synth_ref_inputs(scene, rp, device=dev, cert_family="R", seed=100)
    b.add([inp["cert"][k] for k in range(4)], [inp["warp"][k] for k in range(4)], inp["image"], cams[inp["ref_index"]],
          [cams[q] for q in inp["nbr_indices"]], rng_stream=rp)
descs = eng.upload_descs(b)
params = eng._params(b, cfg, False, 0, 0)
ws_t = eng._ensure_workspace(params)
cur = torch.cuda.current_stream(dev).cuda_stream
def so(reps):
    rc = eng.lib.ldp_debug_launch_stream(C.byref(params), C.c_void_p(descs.data_ptr()), C.c_void_p(ws_t.data_ptr()), C.c_size_t(ws_t.numel()), C.c_void_p(cur), C.c_int(reps))
    assert rc == 0
so(3); torch.cuda.synchronize()
host = (C.c_longlong * (R * 32))()
eng.lib.ldp_debug_read_clocks.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
# zero the clock area, one launch, read
import torch
eng.lib.ldp_debug_read_clocks(C.byref(params), C.c_void_p(ws_t.data_ptr()), host)
before = np.array(host[:12], dtype=np.int64)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); so(1); e1.record(); torch.cuda.synchronize()
eng.lib.ldp_debug_read_clocks(C.byref(params), C.c_void_p(ws_t.data_ptr()), host)
d = np.array(host[:12], dtype=np.int64) - before
names = ["top->wait", "mbar_wait", "stream->A", "control A->B", "publish", "pop_ready", "blocking", "flush", "iters", "in-place", "blocking#", "async pops"]
ncta = 296
print(f"kernel {1e3*e0.elapsed_time(e1):.1f} us")
for n, v in zip(names[:8], d[:8]):
    print(f"{n:14s} {v/ncta/1.965e3:8.2f} us per CTA")
print({n: int(v) for n, v in zip(names[8:], d[8:])})
