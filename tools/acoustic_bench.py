"""Development aid: regular-grid acoustic model (tests/simplewaveequation.py), Gpts/s of the time loop."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from quick_bench import run  # noqa: E402
from opesci_fd_b200 import abi  # noqa: E402
if __name__ == "__main__":
    lib = abi.load_library()
    for n in [int(a) for a in sys.argv[1:]] or [512]:
        for so in (4, 8):
            run("simplewave3d", n, so, 200, False, abi.ARITH_FAST, lib)
            run("simplewave3d", n, so, 200, False, abi.ARITH_REFERENCE, lib)
