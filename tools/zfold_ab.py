"""A/B inside one GPU session: z-fold (z-face ghost loops + z shell slabs in the z-edge tiles of the fused kernel) on / off."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import make_grid
from opesci_fd_b200 import abi
n = int(os.environ.get("AB_N", "1024"))
steps = int(os.environ.get("AB_STEPS", "12"))
lib = abi.load_library()
for rep in range(2):
    for extra, tag in ((0, "pair"), (abi.NO_PAIR, "zfold"), (abi.NO_ZFOLD, "no-zfold")):
        for arith, nm in ((abi.ARITH_FAST, "fast"), (abi.ARITH_REFERENCE, "ref")):
            cfg = dict(kind="eigenwave3d", so=4, grid_size=[n, n, n], dt=0.25 / n, steps=steps, double=False, domain=[1.0, 1.0, 1.0])
            g = make_grid(cfg, flags=arith | abi.HOST_MIRROR_NONE | extra)
            orig = g.build_params
            def wp(orig=orig):
                p, k = orig(); p.warmup_steps = 4; return p, k
            g.build_params = wp
            g.run(library=lib)
            secs, pts, launches = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
            lib.opesci_b200_last_timing(ctypes.byref(secs), ctypes.byref(pts), ctypes.byref(launches))
            l2 = g.convergence_f64()
            kms = (ctypes.c_double * 3)()
            lib.opesci_b200_time_kernels(ctypes.byref(g._arg_grid), 5, kms)
            print("%-9s %-4s fused %.2f ms  ghost %.2f ms  step-loop %.3f ms/step %.2f Gpts/s  launches/step %.1f  l2U %.6e" % (tag, nm, kms[0], kms[2], secs.value / (steps - 4) * 1e3, pts.value * (steps - 4) / secs.value / 1e9, launches.value / (steps - 4.0), l2[0]), flush=True)
            g.free()
