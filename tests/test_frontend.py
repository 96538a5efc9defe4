"""CPU: the host front end (reference API mirror) and the C ABI.

 * FD weights against the table the reference's sympy derivation produces;
 * every float literal of the lowered model against the literals printed in the reference's
   generated sources (tests/golden/literals.json), term by term and in emitted order;
 * the typed C-expression evaluator against C semantics;
 * the CUDA library loads and exports every symbol include/opesci_b200.h declares (no compute).
"""
import ctypes
import json
import os
import re
from fractions import Fraction

import numpy as np
import pytest

from common import GOLDEN, ROOT, load_golden, load_norms, make_grid
from opesci_fd_b200 import abi, cexpr
from opesci_fd_b200.codeprinter import literal, literal_text
from opesci_fd_b200.util import central_weights, staggered_first_weights

LITERALS = json.load(open(os.path.join(GOLDEN, "literals.json")))


def test_staggered_fd_weights_match_reference_table():
    # SURVEY.md 8a, derived by the reference's Taylor-matrix inversion (opesci/util.py:103-118,195-236)
    table = {2: ["1"], 4: ["9/8", "-1/24"], 6: ["75/64", "-25/384", "3/640"],
             8: ["1225/1024", "-245/3072", "49/5120", "-5/7168"],
             10: ["19845/16384", "-735/8192", "567/40960", "-405/229376", "35/294912"],
             12: ["160083/131072", "-12705/131072", "22869/1310720", "-5445/1835008", "847/2359296", "-63/2883584"]}
    for so, coefs in table.items():
        assert staggered_first_weights(so // 2) == [Fraction(c) for c in coefs]
    assert central_weights(2, 2) == [Fraction(-5, 2), Fraction(4, 3), Fraction(-1, 12)]


def test_literal_rounding_rule():
    # 15 significant digits, then a float literal (opesci/codeprinter.py:46-63)
    assert literal_text(9.0 / 8 * 0.2 / 27) == "8.33333333333333e-3"
    assert literal_text(0.225) == "2.25e-1"
    assert literal(1.0 / 3) == np.float32(float("3.33333333333333e-1"))


def _emitted_order_staggered(p, m):
    """Signed literals of the nine interior sums in the order the printer emits them."""
    def window(c, fwd):
        c = [float(np.float32(x)) for x in c[:m]]
        if fwd:
            return c + [-x for x in c[1:]] + [-c[0]]
        return c[1:] + [-x for x in c] + [c[0]]
    out = {}
    for a, name in enumerate(["Txx", "Tyy", "Tzz"]):
        out[name] = sum((window(p.c_stress_normal[a][d], False) for d in range(3)), [])
    for s, name in enumerate(["Txy", "Tyz", "Txz"]):
        out[name] = window(p.c_stress_shear[s][0], True) + window(p.c_stress_shear[s][1], True)
    for a, name in enumerate(["U", "V", "W"]):
        out[name] = sum((window(p.c_velocity[a][d], d == a) for d in range(3)), [])
    return out


def _config(name):
    if os.path.exists(os.path.join(GOLDEN, name + ".npz")):
        return load_golden(name)[0]
    return load_norms()[name]["config"]


@pytest.mark.parametrize("name", sorted(n for n in LITERALS if n.startswith("ew_")))
def test_lowered_staggered_literals_equal_generated_source(name):
    cfg = _config(name)
    grid = make_grid(cfg)
    p, keep = grid.build_params()
    mine = _emitted_order_staggered(p, cfg["so"] // 2)
    for field, lits in LITERALS[name].items():
        ref = [float(np.float32(float(x))) for x in lits]
        assert mine[field] == ref, "%s: literals of %s differ from the generated source" % (name, field)


@pytest.mark.parametrize("name", sorted(n for n in LITERALS if n.startswith("ewh_")))
def test_lowered_heterogeneous_literals_equal_generated_source(name):
    """`read` mode: every term is literal*G*media; own-axis windows of the normal stresses carry a lambda
    and a mu term per offset (literal c and 2c)."""
    cfg = _config(name)
    grid = make_grid(cfg)
    p, keep = grid.build_params()
    m = cfg["so"] // 2
    assert p.hetero == 1

    def window(c, fwd, c2=None):
        c = [float(np.float32(x)) for x in c[:m]]
        seq = c + [-x for x in c[1:]] + [-c[0]] if fwd else c[1:] + [-x for x in c] + [c[0]]
        if c2 is None:
            return seq
        two = [float(np.float32(x)) for x in c2[:m]]
        seq2 = two[1:] + [-x for x in two] + [two[0]]
        return [v for pair in zip(seq, seq2) for v in pair]
    mine = {}
    for a, fname in enumerate(["Txx", "Tyy", "Tzz"]):
        mine[fname] = sum((window(p.h_c[d], False, p.h_c2[d] if d == a else None) for d in range(3)), [])
    for (a, b), fname in zip([(0, 1), (1, 2), (0, 2)], ["Txy", "Tyz", "Txz"]):
        mine[fname] = window(p.h_c[b], True) + window(p.h_c[a], True)
    for a, fname in enumerate(["U", "V", "W"]):
        mine[fname] = sum((window(p.h_c[d], d == a) for d in range(3)), [])
    for field, lits in LITERALS[name].items():
        ref = [float(np.float32(float(x))) for x in lits]
        assert mine[field] == ref, "%s: literals of %s differ from the generated source" % (name, field)


@pytest.mark.parametrize("name", sorted(n for n in LITERALS if n.startswith("sw_")))
def test_lowered_acoustic_literals_equal_generated_source(name):
    cfg = _config(name)
    grid = make_grid(cfg)
    p, keep = grid.build_params()
    m = cfg["so"] // 2
    mine = []
    for d in range(3):
        c = [float(np.float32(p.ac_coef[d][k])) for k in range(m)]
        if any(c):
            mine += c + c
    mine.append(float(np.float32(p.ac_centre)))
    ref = [float(np.float32(float(x))) for x in LITERALS[name]["MAIN_GRID"]]
    assert mine == ref


def test_cexpr_types_like_cxx():
    import math
    v = cexpr.Variables()
    v.scalar("beta", cexpr.FLOAT, 0.7692307692307692)
    v.scalar("mu", cexpr.FLOAT, 1.0530000000000002)
    v.axis("x", cexpr.FLOAT, 0, np.float32(0.1) * np.arange(5, dtype=np.float32))
    v.axis("y", cexpr.FLOAT, 1, np.float32(0.2) * np.arange(4, dtype=np.float32))
    prog = cexpr.compile_expression("(sin(M_PI*y) - 2.0e-3F*x)*cos(1.0e-3F*M_SQRT2*M_PI*sqrt(beta*mu))", v, [5, 4, 3])
    bm = np.float32(0.7692307692307692) * np.float32(1.0530000000000002)            # float*float stays float
    c = math.cos(float(np.float32(1.0e-3)) * 1.41421356237309504880 * 3.14159265358979323846 * math.sqrt(float(bm)))
    for ix in range(5):
        for iy in range(4):
            xs = np.float32(0.1) * np.float32(ix)
            ys = np.float32(0.2) * np.float32(iy)
            want = (math.sin(3.14159265358979323846 * float(ys)) - float(np.float32(2.0e-3) * xs)) * c
            assert prog.evaluate(ix, iy, 0) == want


def test_unsupported_models_raise():
    import eigenwave3d as drv
    # heterogeneous media are fp32 only (the reference's reader is float*), and the files must exist
    h = drv.eigenwave3d((1.0, 1.0, 1.0), (10, 10, 10), 0.002, 0.01, accuracy_order=[2, 4, 4, 4], read=True,
                        double=True, verbose=False)
    with pytest.raises(NotImplementedError):
        h.build_params()
    h = drv.eigenwave3d((1.0, 1.0, 1.0), (10, 10, 10), 0.002, 0.01, accuracy_order=[2, 4, 4, 4], read=True,
                        rho_file="/nonexistent/rho", vp_file="/nonexistent/vp", vs_file="/nonexistent/vs", verbose=False)
    with pytest.raises((IOError, OSError)):
        h.build_params()
    g = drv.eigenwave3d((1.0, 1.0, 1.0), (10, 10, 10), 0.002, 0.01, accuracy_order=[2, 4, 4, 4], verbose=False)
    # any subset of the six faces is accepted (reference: one set_free_surface_boundary call per face)
    g._free_surface.discard((3, 1))
    p, keep = g.build_params()
    assert p.fs_faces == 63 - (1 << 5) and p.free_surface == abi.FS_LEVANDER
    g._free_surface.clear()
    p, keep = g.build_params()
    assert p.fs_faces == 0 and p.free_surface == abi.FS_NONE
    g.set_order([2, 4, 8, 4])
    with pytest.raises(NotImplementedError):
        g.build_params()


def test_cuda_library_exports_the_abi(cuda_lib):
    # loading + symbol lookup only: no compute without a GPU
    # every entry point any header under include/ declares
    hdr = "".join(open(os.path.join(ROOT, "include", h)).read()
                  for h in sorted(os.listdir(os.path.join(ROOT, "include"))) if h.endswith(".h"))
    hdr = re.sub(r"/\*.*?\*/|//[^\n]*", "", hdr, flags=re.S)             # prose in comments is not a declaration
    declared = set(re.findall(r"\b(opesci_\w+)\s*\(", hdr))
    declared -= set(re.findall(r"static\s+inline[^(]*?\b(opesci_\w+)\s*\(", hdr))   # header-only geometry helpers
    assert declared >= {"opesci_execute", "opesci_convergence", "opesci_free", "opesci_b200_configure"}
    for sym in declared:
        assert hasattr(cuda_lib, sym), "libopesci_b200.so does not export %s" % sym
    assert cuda_lib.opesci_b200_is_cuda() == 1
    bad = abi.OpesciB200Params()
    assert cuda_lib.opesci_b200_configure(ctypes.byref(bad)) != 0        # struct_size check, no GPU needed


def test_product_path_never_touches_the_oracle():
    # the package must not import / load anything under oracle/ (oracle = test infrastructure)
    pkg = os.path.join(ROOT, "opesci_fd_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "libopesci_oracle" not in text, f
                assert "import oracle" not in text and "from oracle" not in text, f


def test_generate_writes_model_description(tmp_path):
    import eigenwave3d as drv
    g = drv.eigenwave3d((1.0, 1.0, 1.0), (12, 12, 12), 0.002, 0.01, accuracy_order=[2, 4, 4, 4], verbose=False)
    out = tmp_path / "model.json"
    g.generate(str(out))
    desc = json.loads(out.read_text())
    assert desc["dim"] == [17, 17, 17] and desc["so"] == 4 and desc["free_surface"] == abi.FS_LEVANDER
    assert abs(desc["c_stress_normal"][0][0][0] - 9.0 / 8 * 0.002 * 12 * 1.0) < 1e-6


def test_execute_refuses_a_non_cuda_library(tmp_path):
    # Grid.execute() is the product entry point (reference: opesci/grid.py:85-130): handing it the CPU
    # oracle must fail instead of silently running on the host
    import eigenwave3d as drv
    import __graft_entry__ as ge
    g = drv.eigenwave3d((1.0, 1.0, 1.0), (12, 12, 12), 0.002, 0.01, accuracy_order=[2, 4, 4, 4], verbose=False)
    g.src_lib = ge.build_oracle()
    with pytest.raises(RuntimeError, match="not the CUDA library"):
        g.execute(str(tmp_path / "m.json"))
    g.src_lib = str(tmp_path / "missing.so")
    with pytest.raises(Exception, match="no CPU fallback"):
        g.execute(str(tmp_path / "m.json"))


def test_slab_geometry_properties(cuda_lib, oracle_lib):
    """include/opesci_slab.h through both libraries' `opesci_b200_slab_range` (host arithmetic only, no GPU): for random
    grids, orders and rank counts the stored ranges are the owned ranges plus exactly max(8, need) halo planes per inner
    side, the owned ranges tile the grid, and a decomposition is refused exactly when a slab would be thinner than the halo."""
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=300, deadline=None)
    @given(st.integers(20, 3000), st.sampled_from([2, 4, 6, 8, 10, 12]), st.integers(1, 8))
    def check(gdim, so, nranks):
        m = so // 2
        need = 2 * m + 3 if so == 4 else 2 * m                      # staggered elastic model (opesci_slab_need)
        halo = max(abi.SLAB_HALO, need)
        n_int = gdim - 2 * m
        base, rem = divmod(n_int, nranks)
        prev_hi = 0
        for r in range(nranks):
            got = []
            for lib in (cuda_lib, oracle_lib):
                l0, l1 = ctypes.c_int(-1), ctypes.c_int(-1)
                rc = lib.opesci_b200_slab_range(r, nranks, gdim, so, ctypes.byref(l0), ctypes.byref(l1))
                got.append((rc != 0, l0.value, l1.value))
            assert got[0][0] == got[1][0] == (nranks > 1 and base < halo)
            if got[0][0]:
                continue
            assert got[0] == got[1]
            X0 = m + r * base + min(r, rem)
            X1 = X0 + base + (1 if r < rem else 0)
            own_lo, own_hi = (0 if r == 0 else X0), (gdim if r == nranks - 1 else X1)
            assert own_lo == prev_hi and own_hi > own_lo
            prev_hi = own_hi
            assert got[0][1] == (0 if r == 0 else X0 - halo) and got[0][2] == (gdim if r == nranks - 1 else X1 + halo)
            assert 0 <= got[0][1] and got[0][2] <= gdim
        if not (nranks > 1 and base < halo):
            assert prev_hi == gdim
    check()
