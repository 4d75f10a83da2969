"""Dev tool (build container only): random raw certainty maps, warps and masks (at the map's resolution or another one) through the
LIVE reference's _collect_reference_matches and through the oracle's certainty_prologue: the processed planes must be identical."""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
from oracle import ref_import, densify_oracle as O
from tests.golden.make_prologue_golden import StubMatcher
from tests.golden.make_golden import build_scene
from lichtfeld_densification_plugin_b200 import synth

ref = ref_import.import_reference(full_pipeline=True)
P = ref.pipeline
torch.set_num_threads(1)
rs = np.random.RandomState(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
N = int(sys.argv[2]) if len(sys.argv) > 2 else 30
bad = 0
for t in range(N):
    H, W = int(rs.randint(8, 90)), int(rs.randint(8, 90))
    mh, mw = (H, W) if rs.rand() < 0.4 else (int(rs.randint(5, 120)), int(rs.randint(5, 120)))
    nn = int(rs.randint(1, 4))
    c = dict(H=H, W=W, hm=mh, wm=mw, nn=nn, M=10, fam="T", seed=int(rs.randint(1, 9999)))
    scene = build_scene(c)
    cams = scene.cameras
    floor = float(rs.choice([0.0, 0.2, 0.5, -0.1]))
    raw = (rs.rand(nn, H, W).astype(np.float32) * 1.2 - 0.1).astype(np.float32)
    if t % 4 == 0:
        raw[rs.randint(0, nn), rs.randint(0, H), rs.randint(0, W)] = np.nan
        raw[rs.randint(0, nn), rs.randint(0, H), rs.randint(0, W)] = np.inf
    warp = (rs.rand(nn, H, W, 4).astype(np.float32) * 2.4 - 1.2).astype(np.float32)
    # coordinates exactly on mask-pixel boundaries and half-way points (rounding to nearest even)
    kx = rs.randint(-1, mw + 1, size=(nn, H, W)).astype(np.float64) + 0.5
    pick = rs.rand(nn, H, W) < 0.1
    warp[..., 2][pick] = ((kx + 0.5) / (W / 2) - 1).astype(np.float32)[pick]
    mA = (rs.rand(mh, mw) > 0.3).astype(np.uint8) if rs.rand() < 0.6 else None
    mBs = [(rs.rand(mh, mw) > 0.4).astype(np.uint8) if rs.rand() < 0.7 else None for _ in range(nn)]
    cfg = ref.config.DensePipelineConfig(output_path="/tmp/unused.ply", matches_per_ref=10, certainty_thresh=floor)
    packed = P._PackedReferenceBatch(ref_id=cams[0].uid, ref_path="", imA_np=np.zeros((mh, mw, 3), np.uint8), maskA_np=mA,
                                     wA_cam=cams[0].width, hA_cam=cams[0].height, nn_ids=[cams[1 + k].uid for k in range(nn)],
                                     nn_masks=mBs, nn_arrays=[np.zeros((mh, mw, 3), np.uint8) for _ in range(nn)])
    matcher = StubMatcher([(torch.from_numpy(warp[k]), torch.from_numpy(raw[k].copy())) for k in range(nn)])
    mr, _ = P._collect_reference_matches(packed, matcher, cfg, 0, None)
    for k in range(nn):
        want = mr.cert_list_cpu[k].numpy()
        got = O.certainty_prologue(raw[k], warp[k], mA, mBs[k], floor)
        if not np.array_equal(want, got, equal_nan=True):
            bad += 1
            d = np.argwhere(~((want == got) | (np.isnan(want) & np.isnan(got))))
            print("MISMATCH", t, k, (H, W), (mh, mw), floor, len(d), d[:3].tolist(), want[tuple(d[0])], got[tuple(d[0])])
print(f"{N} random cases, mismatching planes: {bad}")
