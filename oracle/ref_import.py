"""TEST INFRASTRUCTURE ONLY -- live import of the unmodified reference.

Imports ``/root/reference/core`` (the reference plugin's hot-path modules) with the two
stub modules it needs.  This only works in the build container;
``/root/reference`` does not exist on the GPU box, so nothing in ``-m gpu`` tests,
``smoke()`` or ``bench.py`` may call this.  It is used by

* ``tests/golden/make_golden.py`` -- to generate the committed golden vectors, and
* ``tests/test_oracle_vs_reference.py`` -- to pin ``oracle/densify_oracle.py`` (skipped when
  the reference tree is absent).

Nothing under the product package may import this module.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LDP_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "core", "pipeline.py"))


class _NullLog:
    def __getattr__(self, name):
        return lambda *a, **kw: None


def _install_stubs() -> None:
    # `lichtfeld` is the host application's embedded module: only lf.log.* is touched by
    # core/pipeline.py and core/matcher.py.  `pycolmap` is used for annotations only
    # (core/geometry.py:7, core/selection.py) under `from __future__ import annotations`.
    if "lichtfeld" not in sys.modules:
        lf = types.ModuleType("lichtfeld")
        lf.log = _NullLog()
        sys.modules["lichtfeld"] = lf
    if "pycolmap" not in sys.modules:
        sys.modules["pycolmap"] = types.ModuleType("pycolmap")


_cached = {}


def import_reference(full_pipeline: bool = True):
    """Return a namespace with the reference's ``sampling``, ``geometry`` (and ``pipeline``) modules.

    ``full_pipeline=False`` imports only core.sampling / core.geometry (milliseconds);
    ``True`` also imports core.pipeline (several seconds: it pulls the vendored romav2 package,
    but never constructs the network).
    """
    if not reference_available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    key = bool(full_pipeline)
    if key in _cached:
        return _cached[key]
    _install_stubs()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    # The reference's top-level package name is `core`; make sure we are not shadowed.
    mod = sys.modules.get("core")
    if mod is not None and not getattr(mod, "__file__", "").startswith(REFERENCE_ROOT):
        raise RuntimeError("another top-level package named `core` is already imported")
    import importlib

    ns = types.SimpleNamespace()
    ns.sampling = importlib.import_module("core.sampling")
    ns.geometry = importlib.import_module("core.geometry")
    ns.camera_models = importlib.import_module("core.camera_models")
    ns.config = importlib.import_module("core.config")
    ns.writers = importlib.import_module("core.writers")
    ns.image_utils = importlib.import_module("core.image_utils")
    if full_pipeline:
        ns.pipeline = importlib.import_module("core.pipeline")
    _cached[key] = ns
    return ns
