"""Development aid: cost of the per-step .vts output (include/opesci_io.h) on the time loop.
usage: output_bench.py N steps -- prints time-loop seconds and wall seconds without output, every 4th step, every step."""
import ctypes, os, sys, tempfile, time, shutil
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import make_grid  # noqa: E402
from opesci_fd_b200 import abi  # noqa: E402
n, steps = int(sys.argv[1]), int(sys.argv[2])
lib = abi.load_library()
cfg = dict(kind="eigenwave3d", so=4, grid_size=[n, n, n], dt=0.25 / n, steps=steps, double=False, domain=[1.0, 1.0, 1.0])
for every in (0, 4, 1):
    d = tempfile.mkdtemp(prefix="vts_")
    if every:
        lib.opesci_b200_set_output(os.path.join(d, "U_").encode(), 0, every)
    g = make_grid(cfg, flags=abi.ARITH_FAST | abi.HOST_MIRROR_NONE)
    t0 = time.time()
    g.run(library=lib)
    wall = time.time() - t0
    lib.opesci_b200_set_output(None, 0, 0)
    secs, pts, launches = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
    lib.opesci_b200_last_timing(ctypes.byref(secs), ctypes.byref(pts), ctypes.byref(launches))
    files = sorted(os.listdir(d))
    size = sum(os.path.getsize(os.path.join(d, f)) for f in files)
    print("N=%d steps=%d output every %d: time loop %.3f s (%.2f Gpts/s), wall %.2f s, %d files, %.1f MB"
          % (n, steps, every, secs.value, pts.value * steps / secs.value / 1e9, wall, len(files), size / 1e6), flush=True)
    g.free()
    shutil.rmtree(d)
