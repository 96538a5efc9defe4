"""Ad-hoc throughput probe of the heterogeneous (`read`) mode: Gpts/s of the time loop, device-resident
(development aid, not the contract bench).  104 B/point algorithmic (72 + 8 media words)."""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import make_grid  # noqa: E402
from opesci_fd_b200 import abi  # noqa: E402

if __name__ == "__main__":
    lib = abi.load_library()
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 512
    for so in ((4, 8) if len(sys.argv) < 3 else (int(sys.argv[2]),)):
        for arith in (abi.ARITH_REFERENCE, abi.ARITH_FAST):
            steps = 30
            cfg = dict(kind="eigenwave3d_read", so=so, grid_size=[n, n, n], dt=0.2 / n, steps=steps, double=False,
                       domain=[1.0, 1.0, 1.0], seed=1)
            t0 = time.time()
            g = make_grid(cfg, flags=arith | abi.HOST_MIRROR_NONE)
            t1 = time.time()
            orig = g.build_params

            def with_warmup(orig=orig):
                p, k = orig()
                p.warmup_steps = 6
                return p, k
            g.build_params = with_warmup
            g.run(library=lib)
            wall = time.time() - t1
            secs, pts, launches = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
            lib.opesci_b200_last_timing(ctypes.byref(secs), ctypes.byref(pts), ctypes.byref(launches))
            gpts = pts.value * (steps - 6) / secs.value / 1e9
            kms = (ctypes.c_double * 3)()
            lib.opesci_b200_time_kernels(ctypes.byref(g._arg_grid), 5, kms)
            print("hetero n=%d so=%d %s: loop %.3fs  %.2f Gpts/s  %.0f GB/s algorithmic (%.1f%% of 6456)  "
                  "kernels ms %.2f / %.2f / ghost %.2f  media gen %.1fs  execute wall %.1fs"
                  % (n, so, "fast" if arith else "ref ", secs.value, gpts, gpts * 104, gpts * 104 / 64.56, kms[0], kms[1], kms[2],
                     t1 - t0, wall), flush=True)
            g.free()
