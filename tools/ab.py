"""A/B timing of library variants inside ONE gpurun call (boxes differ by several percent)."""
import ctypes, os, sys, glob
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import make_grid
from opesci_fd_b200 import abi
n = int(os.environ.get("AB_N", "1024"))
steps = 8
for rep in range(2):
    for path in sorted(glob.glob(os.path.join(ROOT, "tools", "scratch", "variant_*.so"))):
        lib = abi.load_library(path)
        for arith, nm in ((abi.ARITH_FAST, "fast"), (abi.ARITH_REFERENCE, "ref")):
            cfg = dict(kind="eigenwave3d", so=4, grid_size=[n, n, n], dt=0.25 / n, steps=steps, double=False, domain=[1.0, 1.0, 1.0])
            g = make_grid(cfg, flags=arith | abi.HOST_MIRROR_NONE)
            g.run(library=lib)
            kms = (ctypes.c_double * 3)()
            lib.opesci_b200_time_kernels(ctypes.byref(g._arg_grid), 5, kms)
            secs, pts, launches = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
            lib.opesci_b200_last_timing(ctypes.byref(secs), ctypes.byref(pts), ctypes.byref(launches))
            print("%-28s %-4s fused %.2f ms  ghost %.2f ms  step-loop %.2f Gpts/s" % (os.path.basename(path), nm, kms[0], kms[2], pts.value * steps / secs.value / 1e9), flush=True)
            g.free()
