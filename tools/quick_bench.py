"""Ad-hoc throughput probe (development aid, not the contract bench): Gpts/s of the time loop,
device-resident, for a few grid sizes / orders / arithmetic modes."""
import ctypes
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import make_grid  # noqa: E402
from opesci_fd_b200 import abi  # noqa: E402


def run(kind, n, so, steps, double, arith, lib, extra_flags=0):
    cfg = dict(kind=kind, so=so, grid_size=[n, n, n], dt=0.25 / n, steps=steps, double=double,
               domain=[1.0, 1.0, 1.0])
    g = make_grid(cfg, flags=arith | abi.HOST_MIRROR_NONE | extra_flags)
    t0 = time.time()
    g.run(library=lib)
    wall = time.time() - t0
    secs, pts, launches = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
    lib.opesci_b200_last_timing(ctypes.byref(secs), ctypes.byref(pts), ctypes.byref(launches))
    gpts = pts.value * steps / secs.value / 1e9
    bytes_pt = (72 if kind == "eigenwave3d" else 12) * (2 if double else 1)
    print("%-12s n=%4d so=%2d %s %s steps=%3d  loop %.3fs  %.2f Gpts/s  %.0f GB/s algorithmic (%.1f%% of 6456)  wall %.1fs"
          % (kind, n, so, "f64" if double else "f32", "fast" if arith == abi.ARITH_FAST else "ref ", steps, secs.value,
             gpts, gpts * bytes_pt, gpts * bytes_pt / 64.56, wall), flush=True)
    g.free()
    return gpts


if __name__ == "__main__":
    lib = abi.load_library()
    sizes = [int(a) for a in sys.argv[1:]] or [256, 512]
    for n in sizes:
        steps = 20 if n <= 512 else 10
        for arith in (abi.ARITH_REFERENCE, abi.ARITH_FAST):
            run("eigenwave3d", n, 4, steps, False, arith, lib)
    n = sizes[min(1, len(sizes) - 1)]
    for so in (8, 12):
        for arith in (abi.ARITH_REFERENCE, abi.ARITH_FAST):
            run("eigenwave3d", n, so, 10, False, arith, lib)
    run("eigenwave3d", n, 4, 10, True, abi.ARITH_REFERENCE, lib)
    run("eigenwave3d", n, 4, 10, True, abi.ARITH_FAST, lib)
    run("simplewave3d", n, 4, 30, False, abi.ARITH_REFERENCE, lib)
    run("simplewave3d", n, 4, 30, False, abi.ARITH_FAST, lib)
