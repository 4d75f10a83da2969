"""Dev tool: tests/test_gpu_parity.py::test_randomised_shapes for seeds beyond the 24 the suite runs: python scratch/gpu_fuzz.py LO HI"""
import sys, time
sys.path.insert(0, "/root/repo")
from lichtfeld_densification_plugin_b200.engine import DensifyEngine
from tests import test_gpu_parity as T
eng = DensifyEngine()
lo, hi = int(sys.argv[1]), int(sys.argv[2])
bad = 0
t0 = time.time()
for case in range(lo, hi):
    try:
        T.test_randomised_shapes(eng, case)
    except Exception as exc:      # AssertionError included
        bad += 1
        print("CASE", case, "FAILED:", repr(exc)[:600], flush=True)
print(f"cases {lo}..{hi - 1}: {bad} failed, {time.time() - t0:.0f} s")
