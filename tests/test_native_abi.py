"""CPU: the C-ABI library loads and exports every symbol include/ldp_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

from lichtfeld_densification_plugin_b200 import _native, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    build.build()          # no-op when csrc/libldp_b200.so is up to date
    return _native.load()


def test_header_symbols_exported(lib):
    header = open(os.path.join(ROOT, "include", "ldp_b200.h")).read()
    declared = set(re.findall(r"^\s*(?:int|int64_t|const char\*)\s+(ldp_[a-z_0-9]+)\s*\(", header, flags=re.M))
    assert declared == set(_native.EXPORTS), declared ^ set(_native.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name


def test_abi_version_and_struct_layout(lib):
    header = open(os.path.join(ROOT, "include", "ldp_b200.h")).read()
    assert int(re.search(r"#define LDP_ABI_VERSION (\d+)", header).group(1)) == lib.ldp_abi_version() == _native.LDP_ABI_VERSION
    assert int(re.search(r"#define LDP_MAX_NN (\d+)", header).group(1)) == _native.LDP_MAX_NN
    for which, struct in enumerate((_native.LdpParams, _native.LdpRefDesc, _native.LdpOutputs)):
        assert lib.ldp_struct_size(which) == ctypes.sizeof(struct)
    assert _native.REF_DESC_DTYPE.itemsize == ctypes.sizeof(_native.LdpRefDesc)


def test_host_only_entry_points(lib):
    assert lib.ldp_sel_capacity(10000) == 10000
    assert lib.ldp_sel_capacity(1) == 4
    p = _native.LdpParams()
    p.n_refs, p.H, p.W, p.w_match, p.h_match, p.matches_per_ref, p.border, p.tiles = 46, 512, 512, 512, 512, 10000, 2, 24
    need = ctypes.c_size_t(0)
    assert lib.ldp_workspace_bytes(ctypes.byref(p), ctypes.byref(need)) == 0
    assert 46 * 512 * 512 * 5 < need.value < 46 * 512 * 512 * 16
    p.H = 0
    assert lib.ldp_workspace_bytes(ctypes.byref(p), ctypes.byref(need)) == -1
    assert b"bad" in lib.ldp_last_error_string()


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lichtfeld_densification_plugin_b200.engine import DensifyEngine
    with pytest.raises(_native.NativeLibraryError):
        DensifyEngine()


def test_no_cpu_fallback_around_the_path():
    """The device output contract and pair generation refuse to run without a GPU as well."""
    import numpy as np
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from lichtfeld_densification_plugin_b200 import output
    from lichtfeld_densification_plugin_b200.core import selection
    with pytest.raises(_native.NativeLibraryError):
        output.ply_records(torch.zeros((4, 3)), torch.zeros((4, 3)))
    with pytest.raises(_native.NativeLibraryError):
        output.apply_point_cap(torch.zeros((4, 3)), torch.zeros((4, 3)), torch.zeros((4,)), 2, 0)
    with pytest.raises(_native.NativeLibraryError):
        selection.select_cameras_kcenters(np.zeros((4, 16), np.float32), 2)
    with pytest.raises(_native.NativeLibraryError):
        selection.nearest_neighbors(np.zeros((4, 16), np.float32), 2)
