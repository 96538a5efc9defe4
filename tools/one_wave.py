"""Probe: DRAM traffic of the fused kernel when the whole grid is ONE wave of CTAs (all tiles start together)."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import make_grid
from opesci_fd_b200 import abi
ny = int(sys.argv[1]) if len(sys.argv) > 1 else 91
nx = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
cfg = dict(kind="eigenwave3d", so=4, grid_size=[nx, ny, 1024], dt=0.0002, steps=4, double=False, domain=[1.0, 1.0, 1.0])
lib = abi.load_library()
g = make_grid(cfg, flags=abi.ARITH_FAST | abi.HOST_MIRROR_NONE | abi.NO_CUDA_GRAPH)
g.run(library=lib)
secs, pts, launches = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
lib.opesci_b200_last_timing(ctypes.byref(secs), ctypes.byref(pts), ctypes.byref(launches))
print("points/step %.4g  loop %.4fs  %.2f Gpts/s" % (pts.value, secs.value, pts.value * 4 / secs.value / 1e9))
