"""Development aid: one short run of a given spatial order / precision (for ncu captures)."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from quick_bench import run  # noqa: E402
from opesci_fd_b200 import abi  # noqa: E402
if __name__ == "__main__":
    n, so = int(sys.argv[1]), int(sys.argv[2])
    double = len(sys.argv) > 3 and sys.argv[3] == "f64"
    extra = abi.FORCE_UNFUSED if (len(sys.argv) > 4 and sys.argv[4] == "v1") else 0
    run("eigenwave3d", n, so, 4, double, abi.ARITH_FAST, abi.load_library(), extra)
