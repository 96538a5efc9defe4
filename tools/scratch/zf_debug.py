import ctypes, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import make_grid, fields_of, FIELD_ORDER
from opesci_fd_b200 import abi
import __graft_entry__ as ge
lib = abi.load_library()
ora = abi.bind(ctypes.CDLL(ge.build_oracle()))
for steps in (1, 2):
    cfg = dict(kind="eigenwave3d", so=4, grid_size=[40, 60, 130], dt=0.002, steps=steps, double=False, domain=[1.0, 0.9, 0.8], rho=1.2, vp=1.6, vs=0.8)
    a = make_grid(cfg, flags=abi.ARITH_REFERENCE | abi.HOST_MIRROR_FULL); a.run(library=lib)
    b = make_grid(cfg); b.run(library=ora)
    fa, fb = fields_of(a), fields_of(b)
    print("steps", steps, "dims", fa.shape)
    for k in range(9):
        for lvl in range(2):
            bad = np.argwhere(fa[k, lvl].view(np.int32) != fb[k, lvl].view(np.int32))
            if len(bad):
                print(" field %s level %d: %d bad; x %d..%d y %d..%d z %s" % (FIELD_ORDER[k], lvl, len(bad), bad[:,0].min(), bad[:,0].max(), bad[:,1].min(), bad[:,1].max(), sorted(set(bad[:,2].tolist()))[:12]))
                for q in bad[:3]:
                    print("    ", tuple(q), fa[k, lvl][tuple(q)], fb[k, lvl][tuple(q)])
    a.free(); b.free()
