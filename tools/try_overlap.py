import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import make_grid, fields_of, bits
from opesci_fd_b200 import abi
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 7
cfg = dict(kind="eigenwave3d", so=4, grid_size=[300, 90, 150], dt=0.0005, steps=steps, double=False, domain=[1.0, 1.0, 1.0])
lib = abi.load_library()
for extra, name in ((0, "graph"), (abi.NO_CUDA_GRAPH, "nograph")):
    a = make_grid(cfg, flags=abi.ARITH_REFERENCE | abi.HOST_MIRROR_FULL | abi.OVERLAP | extra)
    a.run(library=lib)
    b = make_grid(cfg, flags=abi.ARITH_REFERENCE | abi.HOST_MIRROR_FULL | abi.FORCE_UNFUSED)
    b.run(library=lib)
    fa, fb = fields_of(a), fields_of(b)
    print(name, "differing cells per field:", [int((bits(fa[k]) != bits(fb[k])).sum()) for k in range(9)])
    a.free(); b.free()
