"""One-off GPU diagnostic (not a test): run a golden case + a fast-size case and print parity details."""
import sys, time, json
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from tests.helpers import load_golden, GOLDEN_CASES
from tests import gpu_harness as G
from lichtfeld_densification_plugin_b200.engine import DensifyEngine
eng = DensifyEngine()
for name in GOLDEN_CASES:
    try:
        c, scene, inp, z = load_golden(name)
        U = np.random.RandomState(int(z["mt_seed"])).random_sample(3 * c["M"] + 64)
        s = z["weight_sum"]
        res = G.run_oracle_ref(scene, inp, c, uniforms=None if c["no_filter"] else U, s_override=s, collect_debug=True)
        g = G.run_gpu(eng, scene, [inp], G.path_cfg(c), uniforms=U[None, :], weight_sums=[s], collect_debug=True)
        rep = G.compare_ref(g, 0, res, c, scene)
        print(name, "status", g.status, "S gpu/ref", g.sel_idx[0].size, res.sel_idx.size, "used", g.uniforms_used, res.taps.get("uniforms_used"), "rounds", g.rounds, res.taps.get("rounds"))
        print("   ", rep)
        if not rep.sel_exact:
            a, b = g.sel_idx[0], res.sel_idx
            print("    only_gpu", np.setdiff1d(a, b)[:20], "only_ref", np.setdiff1d(b, a)[:20])
    except Exception as e:
        import traceback; traceback.print_exc()
