"""Golden vectors for the certainty post-processing step that precedes the path, made from the
LIVE, UNMODIFIED reference: ``core.pipeline._collect_reference_matches`` (core/pipeline.py:385-460) driven by a stub
matcher that returns prepared (warp, certainty) maps, followed by ``_triangulate_ref`` on its output.

Run in the build container only (needs /root/reference):

    python tests/golden/make_prologue_golden.py
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402
from lichtfeld_densification_plugin_b200 import synth  # noqa: E402
from tests.golden.make_golden import build_scene  # noqa: E402

CASES = {
    # maps 96^2 from 64^2 masks (the 'precise' situation: map resolution != match resolution), both masks
    "masks_resized": dict(H=96, W=96, hm=64, wm=64, nn=3, M=3000, fam="T", seed=21, maskA=True, maskB=[True, False, True]),
    # same resolution, reference mask only
    "maskA_only": dict(H=64, W=80, hm=64, wm=80, nn=2, M=1500, fam="T", seed=22, maskA=True, maskB=[False, False]),
    # no masks: the floor clamp alone
    "floor_only": dict(H=64, W=64, hm=64, wm=64, nn=2, M=1200, fam="T", seed=23, maskA=False, maskB=[False, False]),
}
CERTAINTY_THRESH = 0.2


class StubMatcher:
    def __init__(self, pairs):
        self.pairs = pairs

    def match_grids_batch(self, imA, nn_images):
        assert len(nn_images) == len(self.pairs)
        return list(self.pairs)


def make_masks(c, rs):
    hm, wm = c["hm"], c["wm"]
    yy, xx = np.mgrid[0:hm, 0:wm]
    mA = (((xx - 0.45 * wm) ** 2 + (yy - 0.55 * hm) ** 2) < (0.42 * min(hm, wm)) ** 2).astype(np.uint8) if c["maskA"] else None
    mBs = []
    for k, on in enumerate(c["maskB"]):
        if not on:
            mBs.append(None)
            continue
        m = ((xx * (1 + k) + yy * 2) % 23 > 3).astype(np.uint8)          # stripes: many nearest-neighbour decisions
        m[: hm // 6] = 0
        mBs.append(m)
    return mA, mBs


def raw_inputs(c):
    scene = build_scene(c)
    inp = synth.synth_ref_inputs(scene, 0, cert_family=c["fam"], seed=c["seed"])
    rs = np.random.RandomState(c["seed"])
    raw = inp["cert"].numpy().copy()
    raw -= np.float32(0.12)                                  # part of the map falls below the floor
    warp = inp["warp"].numpy().copy()
    # some matches leave the neighbour image, some land exactly between two mask pixels
    H, W = c["H"], c["W"]
    k = rs.randint(-1, W + 1, size=(H, W)).astype(np.float64) + 0.5
    edge = ((k + 0.5) / (W / 2) - 1).astype(np.float32)
    pick = rs.rand(H, W) < 0.02
    warp[0][..., 2][pick] = edge[pick]
    out = rs.rand(H, W) < 0.01
    warp[0][..., 3][out] = np.float32(1.3)
    return scene, inp, raw, warp


def main() -> None:
    ref = ref_import.import_reference(full_pipeline=True)
    P = ref.pipeline
    torch.set_num_threads(1)
    for name, c in CASES.items():
        scene, inp, raw, warp = raw_inputs(c)
        rs = np.random.RandomState(c["seed"] + 100)
        mA, mBs = make_masks(c, rs)
        cams = scene.cameras
        ri, nb = inp["ref_index"], inp["nbr_indices"]
        cfg = ref.config.DensePipelineConfig(output_path="/tmp/unused.ply", matches_per_ref=c["M"],
                                             certainty_thresh=CERTAINTY_THRESH)
        nn_arrays = [np.zeros((c["hm"], c["wm"], 3), dtype=np.uint8) for _ in nb]
        packed = P._PackedReferenceBatch(ref_id=cams[ri].uid, ref_path="", imA_np=inp["image"].numpy(), maskA_np=mA,
                                         wA_cam=cams[ri].width, hA_cam=cams[ri].height,
                                         nn_ids=[cams[j].uid for j in nb], nn_masks=mBs, nn_arrays=nn_arrays)
        matcher = StubMatcher([(torch.from_numpy(warp[k]), torch.from_numpy(raw[k])) for k in range(len(nb))])
        mr, counter = P._collect_reference_matches(packed, matcher, cfg, 0, None)
        assert mr is not None and counter == len(nb)
        ctx = P._TriangulationContext(cameras=P._build_camera_lookup(cams), config=cfg, matcher_sample_cap=0.9,
                                      w_match=c["wm"], h_match=c["hm"])
        np.random.seed(c["seed"])
        out = P._triangulate_ref(mr, ctx, collect_debug_matches=True)
        assert out is not None, name
        cert_post = np.stack([t.numpy() for t in mr.cert_list_cpu])
        best = torch.max(torch.stack(mr.cert_list_cpu), dim=0).values
        capped = torch.clamp(best.clone(), max=0.9)
        Hh, Ww = capped.shape
        yy, xx = torch.meshgrid(torch.arange(Hh), torch.arange(Ww), indexing="ij")
        inside = (xx >= 2) & (xx <= Ww - 3) & (yy >= 2) & (yy <= Hh - 3)
        s = np.float32((capped * inside.float()).reshape(-1).sum().item())
        np.random.seed(c["seed"])
        sel_idx = ref.sampling.select_samples_with_coverage(best, c["M"], cap=0.9, border=2, tiles=24)
        payload = dict(case=json.dumps({**c, "certainty_thresh": CERTAINTY_THRESH, "numpy": np.__version__, "torch": torch.__version__}),
                       ref_index=np.int64(ri), nbr_indices=np.asarray(nb, dtype=np.int64), mt_seed=np.int64(c["seed"]),
                       raw_cert=raw, warp=warp, image=inp["image"].numpy(), cert_post=cert_post, weight_sum=s,
                       sel_idx=sel_idx.astype(np.int64), xyz=out.xyz, rgb=out.rgb, err=out.err,
                       dbg_uids=np.asarray(list(out.debug_matches_by_nbr.keys()), dtype=np.int64),
                       has_maskA=np.int64(mA is not None), has_maskB=np.asarray([m is not None for m in mBs], dtype=np.int64))
        if mA is not None:
            payload["maskA"] = mA
        for k, m in enumerate(mBs):
            if m is not None:
                payload[f"maskB_{k}"] = m
        for uid, m in out.debug_matches_by_nbr.items():
            payload[f"dbg_matches_{uid}"] = m
            payload[f"dbg_cert_{uid}"] = out.debug_cert_by_nbr[uid]
        path = os.path.join(HERE, "prologue", f"{name}.npz")
        np.savez_compressed(path, **payload)
        zero = float((cert_post == 0).mean())
        print(f"{name}: S={sel_idx.size} K={out.xyz.shape[0]} masked={zero:.2f} s={s!r} -> {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
