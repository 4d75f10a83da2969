"""Launch the first stage of the path alone a few times (for ncu)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
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
rc = eng.lib.ldp_debug_launch_stream(C.byref(params), C.c_void_p(descs.data_ptr()), C.c_void_p(ws_t.data_ptr()), C.c_size_t(ws_t.numel()), C.c_void_p(cur), C.c_int(int(sys.argv[1]) if len(sys.argv) > 1 else 4))
assert rc == 0
torch.cuda.synchronize()
print("ok")
