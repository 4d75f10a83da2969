"""Dev tool: one seed of test_randomised_shapes, with the differing sampled indices explained."""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
from lichtfeld_densification_plugin_b200 import synth
from lichtfeld_densification_plugin_b200.engine import DensifyEngine
from tests import gpu_harness as G
case = int(sys.argv[1])
eng = DensifyEngine()
rs = np.random.RandomState(1000 + case)
H = int(rs.choice([40, 57, 64, 96, 128, 150, 33, 200])); W = int(rs.choice([48, 61, 64, 100, 128, 256, 52, 36]))
hm = int(rs.choice([H, max(16, H // 2), H + 7])); wm = int(rs.choice([W, max(16, W // 2), W + 5]))
nn = int(rs.randint(1, 9)) if case >= 10 else int(rs.randint(1, 6))
M = int(rs.choice([50, 400, 1500, min(4000, H * W // 3)])) if case < 10 else int(rs.choice([8, 123, 900, 2500, H * W // 2]))
fam = "T" if case % 3 else "R"
no_filter = bool(case in (7, 15, 21))
scene = synth.make_scene(10, "turbo", ref_fraction=0.2, nn=nn)
scene.H, scene.W, scene.h_match, scene.w_match = H, W, hm, wm
c = dict(M=M, no_filter=no_filter, wm=wm, hm=hm, sampson=float(rs.choice([5.0, 0.0])), parallax=float(rs.choice([0.5, 0.0])))
inputs = [synth.synth_ref_inputs(scene, rp, cert_family=fam, seed=200 + case) for rp in range(scene.n_refs)]
U = np.stack([np.random.RandomState(case * 10 + r).random_sample(3 * M + 64) for r in range(len(inputs))])
print("case", case, dict(H=H, W=W, hm=hm, wm=wm, nn=nn, M=M, fam=fam), c, "tile", max(1, W // 24))
ress = []
for r, inp in enumerate(inputs):
    try:
        ress.append(G.run_oracle_ref(scene, inp, c, uniforms=U[r]))
    except Exception as exc:
        print("oracle raised for ref", r, repr(exc)); ress.append(None)
try:
    g = G.run_gpu(eng, scene, inputs, G.path_cfg(c), uniforms=U, weight_sums=[res.taps["s"] if res is not None else 0.0 for res in ress])
except Exception as exc:
    print("GPU raised:", repr(exc)); sys.exit(0)
for r, res in enumerate(ress):
    if res is None:
        print("ref", r, "oracle None; gpu status", g.status[r], "n", g.sel_idx[r].size); continue
    a, b = g.sel_idx[r], res.sel_idx
    p = res.taps["p"]; main = res.taps["idx_main"]
    only_g = np.setdiff1d(a, b); only_o = np.setdiff1d(b, a)
    order = np.argsort(-p, kind="stable"); rank = np.empty_like(order); rank[order] = np.arange(order.size)
    print(f"ref {r}: S gpu {a.size} oracle {b.size}, main {main.size}, budget {max(1, M - main.size)}, positive pixels {(p > 0).sum()}, uniforms gpu {g.uniforms_used[r]} oracle {res.taps['uniforms_used']}, rounds {g.rounds[r]} {res.taps['rounds']}")
    print("   only gpu   :", [(int(i), float(p[i]), int(rank[i]), bool(i in set(main.tolist()))) for i in only_g[:8]])
    print("   only oracle:", [(int(i), float(p[i]), int(rank[i]), bool(i in set(main.tolist()))) for i in only_o[:8]])
