"""Golden vectors for the post-path reducers, from the LIVE, UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python tests/golden/make_output_golden.py

* ``_apply_point_cap`` lives in the reference's top-level ``densify.py``, which cannot be imported outside the host
  application (relative imports of the plugin package, ``lichtfeld``, ``pycolmap`` at import time): the function's own
  source is taken from that file with ``ast`` and executed unchanged in a namespace that holds numpy only.
* ``_build_filtered_match_preview`` is imported from ``core.pipeline`` (oracle/ref_import.py).

Inputs are regenerated from the stored seeds by the tests; only the reference's outputs are stored.
"""
from __future__ import annotations

import ast
import os
import sys
from typing import Optional, Tuple  # noqa: F401  (names used by the extracted source)

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import ref_import  # noqa: E402


def extract_function(path: str, name: str):
    src = open(path).read()
    tree = ast.parse(src)
    for node in tree.body:
        if isinstance(node, ast.FunctionDef) and node.name == name:
            code = ast.get_source_segment(src, node)
            ns = {"np": np, "Tuple": Tuple, "Optional": Optional}
            exec(compile("from __future__ import annotations\n" + code, path, "exec"), ns)
            return ns[name]
    raise RuntimeError(f"{name} not found in {path}")


def cap_inputs(n: int, seed: int):
    rs = np.random.RandomState(seed)
    return (rs.standard_normal((n, 3)).astype(np.float32), rs.random_sample((n, 3)).astype(np.float32),
            rs.random_sample((n,)).astype(np.float32))


def preview_inputs(k: int, seed: int):
    rs = np.random.RandomState(seed)
    return (rs.random_sample((k, 4)).astype(np.float32) * 511.0), rs.random_sample((k,)).astype(np.float32)


CAP_CASES = [dict(n=5000, max_points=1234, seed=7, data_seed=21), dict(n=300, max_points=299, seed=0, data_seed=22),
             dict(n=100, max_points=100, seed=3, data_seed=23), dict(n=100, max_points=0, seed=3, data_seed=24)]
PREVIEW_CASES = [dict(k=4000, max_matches=1500, ref_id=5, nbr_id=9, data_seed=31),
                 dict(k=900, max_matches=1500, ref_id=2 ** 33 + 17, nbr_id=123456789, data_seed=32),
                 dict(k=2500, max_matches=700, ref_id=2 ** 33 + 17, nbr_id=123456789, data_seed=33)]


def main() -> None:
    apply_point_cap = extract_function(os.path.join(ref_import.REFERENCE_ROOT, "densify.py"), "_apply_point_cap")
    ref = ref_import.import_reference(full_pipeline=True)
    out = {}
    for i, c in enumerate(CAP_CASES):
        xyz, rgb, err = cap_inputs(c["n"], c["data_seed"])
        a, b, e = apply_point_cap(xyz, rgb, err, c["max_points"], c["seed"])
        out[f"cap{i}_xyz"], out[f"cap{i}_rgb"], out[f"cap{i}_err"] = a, b, e
        out[f"cap{i}_case"] = np.array([c["n"], c["max_points"], c["seed"], c["data_seed"]], dtype=np.int64)
    img = np.zeros((4, 4, 3), np.uint8)
    for i, c in enumerate(PREVIEW_CASES):
        m, cn = preview_inputs(c["k"], c["data_seed"])
        pv = ref.pipeline._build_filtered_match_preview(img, img, m, cn, c["ref_id"], c["nbr_id"], "a", "b", 0, 1, c["k"],
                                                        max_matches=c["max_matches"])
        out[f"pv{i}_matches"], out[f"pv{i}_cert"] = pv.matches, pv.cert_norm
        out[f"pv{i}_case"] = np.array([c["k"], c["max_matches"], c["ref_id"], c["nbr_id"], c["data_seed"]], dtype=np.int64)
    out["numpy_version"] = np.array(np.__version__)
    path = os.path.join(HERE, "output_reducers.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
