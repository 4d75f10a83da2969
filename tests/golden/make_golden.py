"""Generate the golden vectors under tests/golden/ from the LIVE, UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

For every case the synthetic inputs are built with ``lichtfeld_densification_plugin_b200.synth``
(CPU, seeded), handed to the reference's own ``core.pipeline._triangulate_ref`` after
``np.random.seed(case seed)``, and the outputs are frozen into ``<case>.npz`` together with the
inputs (small cases) or an input checksum (config-1 case), the MT19937 seed, the f32 weight sum
``s`` the reference's torch-CPU reduction produced on this box, and the library versions.
The reference has no tests or fixtures of its own for this path -- these
files are what pins the oracle and, through it, the CUDA path.
"""
from __future__ import annotations

import hashlib
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
from lichtfeld_densification_plugin_b200.core.camera_models import CameraRecord  # noqa: E402

CASES = {
    # name: dict(H, W, h_match, w_match, nn, M, family, no_filter, debug, store_inputs, seed, filters)
    "tiny_T_budget_binds": dict(H=48, W=64, hm=48, wm=64, nn=2, M=600, fam="T", no_filter=False, seed=11, store=True),
    "small_T_all_bins": dict(H=96, W=96, hm=96, wm=96, nn=3, M=4000, fam="T", no_filter=False, seed=12, store=True),
    "small_R_debug": dict(H=96, W=96, hm=64, wm=64, nn=3, M=3000, fam="R", no_filter=False, seed=13, store=True),
    "small_T_nofilter": dict(H=64, W=64, hm=64, wm=64, nn=3, M=500, fam="T", no_filter=True, seed=14, store=True),
    "small_T_no_sampson_no_parallax": dict(H=64, W=80, hm=64, wm=80, nn=2, M=1500, fam="T", no_filter=False, seed=15,
                                           store=True, sampson=0.0, parallax=0.0),
    "config1_turbo_1nn": dict(H=320, W=320, hm=320, wm=320, nn=1, M=10000, fam="T", no_filter=False, seed=16, store=False),
}


def build_scene(c) -> synth.SynthScene:
    cams = synth.make_orbit_cameras(9)
    centres = torch.from_numpy(np.stack([cam.C for cam in cams]))
    d = torch.cdist(centres, centres)
    d.fill_diagonal_(float("inf"))
    nn_table = torch.topk(d, c["nn"], largest=False, dim=1).indices.numpy()
    return synth.SynthScene(cameras=cams, refs_local=[4], nn_table=nn_table, H=c["H"], W=c["W"],
                            h_match=c["hm"], w_match=c["wm"], nn=c["nn"])


def input_digest(inp) -> str:
    h = hashlib.sha256()
    for k in ("cert", "warp", "image"):
        h.update(np.ascontiguousarray(inp[k].numpy()).tobytes())
    return h.hexdigest()


def main() -> None:
    ref = ref_import.import_reference(full_pipeline=True)
    P = ref.pipeline
    torch.set_num_threads(1)
    meta = {"numpy": np.__version__, "torch": torch.__version__, "torch_threads": torch.get_num_threads(),
            "reference": "shadygm/Lichtfeld-Densification-Plugin v0.8.3 (pyproject.toml:6)"}
    for name, c in CASES.items():
        scene = build_scene(c)
        inp = synth.synth_ref_inputs(scene, 0, cert_family=c["fam"], seed=c["seed"])
        cams = scene.cameras
        ri, nb = inp["ref_index"], inp["nbr_indices"]
        cfg = ref.config.DensePipelineConfig(output_path="/tmp/unused.ply", matches_per_ref=c["M"],
                                             no_filter=c["no_filter"],
                                             sampson_thresh=c.get("sampson", 5.0),
                                             min_parallax_deg=c.get("parallax", 0.5))
        ctx = P._TriangulationContext(cameras=P._build_camera_lookup(cams), config=cfg, matcher_sample_cap=0.9,
                                      w_match=c["wm"], h_match=c["hm"])
        packed = P._PackedReferenceBatch(ref_id=cams[ri].uid, ref_path="", imA_np=inp["image"].numpy(), maskA_np=None,
                                         wA_cam=cams[ri].width, hA_cam=cams[ri].height,
                                         nn_ids=[cams[j].uid for j in nb], nn_masks=[None] * len(nb),
                                         nn_arrays=[None] * len(nb))
        mr = P._MatchedReference(packed=packed, warp_list_cpu=[inp["warp"][k] for k in range(len(nb))],
                                 cert_list_cpu=[inp["cert"][k] for k in range(len(nb))],
                                 pair_index_by_nbr={}, image_by_nbr={})
        np.random.seed(c["seed"])
        out = P._triangulate_ref(mr, ctx, collect_debug_matches=True)
        assert out is not None, name
        # the sampler's own output and the f32 weight sum as the reference computes them on this box
        best_cert = torch.max(torch.stack([inp["cert"][k] for k in range(len(nb))]), dim=0).values
        np.random.seed(c["seed"])
        sel_idx = ref.sampling.select_samples_with_coverage(best_cert, c["M"], cap=0.9, border=2, tiles=24,
                                                            no_filter=c["no_filter"])
        capped = torch.clamp(best_cert.clone(), max=0.9)
        Hh, Ww = capped.shape
        yy, xx = torch.meshgrid(torch.arange(Hh), torch.arange(Ww), indexing="ij")
        inside = (xx >= 2) & (xx <= Ww - 3) & (yy >= 2) & (yy <= Hh - 3)
        s = np.float32((capped * inside.float()).reshape(-1).sum().item())
        payload = dict(
            case=json.dumps({**c, **meta}),
            ref_index=np.int64(ri), nbr_indices=np.asarray(nb, dtype=np.int64), mt_seed=np.int64(c["seed"]),
            weight_sum=s, sel_idx=sel_idx.astype(np.int64),
            xyz=out.xyz, rgb=out.rgb, err=out.err,
            dbg_uids=np.asarray(list(out.debug_matches_by_nbr.keys()), dtype=np.int64),
            input_sha256=input_digest(inp),
        )
        for uid, m in out.debug_matches_by_nbr.items():
            payload[f"dbg_matches_{uid}"] = m
            payload[f"dbg_cert_{uid}"] = out.debug_cert_by_nbr[uid]
        if c["store"]:
            payload.update(cert=inp["cert"].numpy(), warp=inp["warp"].numpy(), image=inp["image"].numpy())
        path = os.path.join(HERE, f"{name}.npz")
        np.savez_compressed(path, **payload)
        print(f"{name}: S={sel_idx.size} K={out.xyz.shape[0]} s={s!r} -> {os.path.getsize(path) / 1024:.0f} KiB")

    # writers + colour quantisation (reference core/writers.py:15-46, core/image_utils.py:24-26)
    rng = np.random.RandomState(5)
    xyz = rng.standard_normal((257, 3)).astype(np.float32) * 3
    rgb = rng.random_sample((257, 3)).astype(np.float32)
    rgb[:8] = np.array([0.5 / 255, 1.5 / 255, 2.5 / 255, 0.0, 1.0, 1.2, -0.1, 254.5 / 255], dtype=np.float32)[:, None]
    err = rng.random_sample(257).astype(np.float32)
    u8 = ref.image_utils.to_uint8_rgb(rgb)
    tmp = "/tmp/_ldp_golden"
    os.makedirs(tmp, exist_ok=True)
    ref.writers.write_ply(os.path.join(tmp, "a.ply"), xyz, u8)
    ref.writers.write_points3D_bin(os.path.join(tmp, "a.bin"), xyz, u8, err)
    ref.writers.write_points3D_bin(os.path.join(tmp, "b.bin"), xyz, u8, None)
    np.savez_compressed(os.path.join(HERE, "writers.npz"), xyz=xyz, rgb=rgb, err=err, rgb_u8=u8,
                        ply=np.frombuffer(open(os.path.join(tmp, "a.ply"), "rb").read(), dtype=np.uint8),
                        bin=np.frombuffer(open(os.path.join(tmp, "a.bin"), "rb").read(), dtype=np.uint8),
                        bin_noerr=np.frombuffer(open(os.path.join(tmp, "b.bin"), "rb").read(), dtype=np.uint8))
    print("writers.npz written")


if __name__ == "__main__":
    main()
