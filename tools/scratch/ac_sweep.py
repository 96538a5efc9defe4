import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests")); sys.path.insert(0, os.path.join(ROOT, "tools"))
from quick_bench import run
from opesci_fd_b200 import abi
lib = abi.load_library()
run("simplewave3d", 512, 4, 200, False, abi.ARITH_FAST, lib)
run("simplewave3d", 512, 4, 200, False, abi.ARITH_REFERENCE, lib)
