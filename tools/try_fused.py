import sys, os, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from common import make_grid, fields_of, bits
from opesci_fd_b200 import abi
import __graft_entry__ as ge
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
so = int(sys.argv[2]) if len(sys.argv) > 2 else 4
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
cfg = dict(kind="eigenwave3d", so=so, grid_size=[n, n + 3, n + 7], dt=0.1 / n, steps=steps, double=False, domain=[1.0, 1.0, 1.0])
lib = abi.load_library()
a = make_grid(cfg, flags=abi.ARITH_REFERENCE | abi.HOST_MIRROR_FULL | abi.NO_CUDA_GRAPH)
a.run(library=lib)
print("fused run ok")
b = make_grid(cfg, flags=abi.ARITH_REFERENCE | abi.HOST_MIRROR_FULL | abi.FORCE_UNFUSED)
b.run(library=lib)
fa, fb = fields_of(a), fields_of(b)
names = ["U", "V", "W", "Txx", "Tyy", "Tzz", "Txy", "Tyz", "Txz"]
for k in range(9):
    bad = np.argwhere(bits(fa[k]) != bits(fb[k]))
    print(names[k], "differing cells:", len(bad), "first:", bad[:3].tolist() if len(bad) else "")
