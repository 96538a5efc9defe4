"""Development aid: one short run of a given spatial order / precision (for ncu captures)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from quick_bench import run  # noqa: E402
from opesci_fd_b200 import abi  # noqa: E402
if __name__ == "__main__":
    n, so = int(sys.argv[1]), int(sys.argv[2])
    double = len(sys.argv) > 3 and sys.argv[3] == "f64"
    mode = sys.argv[4] if len(sys.argv) > 4 else ""
    extra = {"v1": abi.FORCE_UNFUSED, "tiled": abi.FORCE_TILED}.get(mode, 0)
    steps = int(sys.argv[5]) if len(sys.argv) > 5 else 4
    run("eigenwave3d", n, so, steps, double, abi.ARITH_FAST, abi.load_library(os.environ.get('OPESCI_LIB')), extra)
