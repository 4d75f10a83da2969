"""B200-native post-matching densification hot path (sample -> triangulate -> filter -> colour).

Drop-in for the reference plugin's ``core.pipeline._triangulate_ref`` and its callees; see DESIGN.md.
The work runs in hand-written sm_100a CUDA kernels behind ``csrc/libldp_b200.so`` (C ABI declared in
``include/ldp_b200.h``); there is no CPU fallback -- calling an op without the library raises.
"""
__version__ = "0.1.0"
