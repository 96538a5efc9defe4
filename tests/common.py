"""Helpers shared by the tests: build a front-end grid from a reference configuration."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
REF_MANIFEST = os.path.join(ROOT, "oracle", "_ref", "manifest.json")

FIELD_ORDER = ["U", "V", "W", "Txx", "Tyy", "Tzz", "Txy", "Tyz", "Txz"]


def golden_names(prefix=""):
    # io_*.npz are the vectors of include/opesci_io.h (tests/test_io.py), not field fixtures
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.endswith(".npz") and f.startswith(prefix) and not f.startswith("io_"))


def load_golden(name):
    z = np.load(os.path.join(GOLDEN, name + ".npz"))
    return json.loads(str(z["config"])), z["fields"], z["l2"]


def load_norms():
    return json.load(open(os.path.join(GOLDEN, "norms.json")))


def make_grid(cfg, flags=None):
    """cfg: dict with kind, so, grid_size, dt, steps, double, domain[, rho, vp, vs]."""
    so = cfg["so"]
    order = [2, so, so, so]
    if cfg["kind"] == "eigenwave3d_read":
        # heterogeneous `read` mode: the same synthetic medium the patched reference read from files
        import eigenwave3d as drv
        from opesci_fd_b200.util import synthetic_media
        g = drv.eigenwave3d(tuple(cfg["domain"]), tuple(cfg["grid_size"]), cfg["dt"], cfg["dt"] * cfg["steps"],
                            accuracy_order=order, o_converge=False, read=True, rho_file="rho", vp_file="vp",
                            vs_file="vs", verbose=False)
        g.set_media_arrays(*synthetic_media([d.value for d in g.dim], cfg["seed"]))
    elif cfg["kind"] == "eigenwave3d":
        import eigenwave3d as drv
        g = drv.eigenwave3d(tuple(cfg["domain"]), tuple(cfg["grid_size"]), cfg["dt"], cfg["dt"] * cfg["steps"],
                            accuracy_order=order, o_converge=True, double=cfg["double"],
                            rho=cfg.get("rho", 1.0), vp=cfg.get("vp", 1.0), vs=cfg.get("vs", 0.5), verbose=False)
        if "faces" in cfg:   # free surfaces on a subset of the faces (the driver sets all six)
            g._free_surface = {tuple(f) for f in cfg["faces"]}
    elif cfg["kind"] == "regular_generic":
        # PDE systems outside the fixed-function kernels (tests/generic_pdes.py): NVRTC path
        import generic_pdes
        import opesci_fd_b200
        g = generic_pdes.build(opesci_fd_b200, cfg["pde"], tuple(cfg["domain"]), tuple(cfg["grid_size"]), cfg["dt"],
                               cfg["dt"] * cfg["steps"], order, double=cfg["double"])
    else:
        import simplewaveequation as drv
        g = drv.simplewave3d(tuple(cfg["domain"]), tuple(cfg["grid_size"]), cfg["dt"], cfg["dt"] * cfg["steps"],
                             accuracy_order=order, o_converge=True, double=cfg["double"], verbose=False)
    g.ntsteps.value = cfg["steps"]   # never let tmax/dt rounding decide the step count
    if flags is not None:
        g.b200_flags = flags
    return g


def fields_of(grid):
    return np.stack([grid.field_array(k).copy() for k in range(len(grid.fields))])


def bits(a):
    return a.view(np.int64 if a.dtype == np.float64 else np.int32)


def rel_l2(a, b):
    a = a.astype(np.float64)
    b = b.astype(np.float64)
    den = np.sqrt((b * b).sum())
    return float(np.sqrt(((a - b) ** 2).sum()) / den) if den > 0 else float(np.sqrt(((a - b) ** 2).sum()))
