"""Dev tool: rebuild the library with extra nvcc flags: python scratch/build_with.py -DK2_MIN_BLOCKS=7 ..."""
import sys
sys.path.insert(0, "/root/repo")
from lichtfeld_densification_plugin_b200 import build as B
B.NVCC_FLAGS.extend(sys.argv[1:])
print(B.build(force=True))
