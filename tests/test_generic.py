"""PDE systems outside the fixed-function kernels: RegularGrid.solve_fd derives the update like the
reference (opesci/regulargrid.py:230-270), prints the same expressions, and the library compiles them for sm_100a with
NVRTC.  Pinned on the reference itself: tests/golden/gen_*.npz and generic_kernels.json come from the reference's own
generator run on the same PDE definitions (tests/generic_pdes.py through oracle/refgen/make_ref.py)."""
import json
import os

import numpy as np
import pytest

from common import GOLDEN, bits, fields_of, load_golden, load_norms, make_grid
from opesci_fd_b200 import abi

KERNELS = json.load(open(os.path.join(GOLDEN, "generic_kernels.json")))


@pytest.mark.parametrize("name", sorted(KERNELS))
def test_printed_kernels_equal_the_reference_generators_text(name):
    """CPU: the assignments of the time loop and of the second initialisation, character for character."""
    entry = KERNELS[name]
    g = make_grid(entry["config"])
    assert g.generic
    step, init2 = g.generic_kernel_text()
    assert step == entry["step"]
    assert init2 == entry["init2"]
    p, keep = g.build_params()
    assert p.kind == abi.KIND_REGULAR_GENERIC and p.nfields == len(entry["config"]["fields"]) and p.nlevels == 3
    src = g.generic_source
    assert 'extern "C" __global__' in src and "opesci_generic_step" in src and "opesci_generic_init2" in src
    for line in step + init2:
        assert line in src


def test_acoustic_form_still_takes_the_fixed_function_kernels():
    cfg = load_norms()["sw_default_so4_f32"]["config"]
    g = make_grid(cfg)
    assert not g.generic and g.axis_weights is not None


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(n for n in KERNELS if "_mid_" in n))
def test_generic_norms_reproduce_the_reference(name, cuda_lib):
    """64^3 x 60 steps: the reference's printed L2 output, digit for digit, by the library itself (OPESCI_L2_REFERENCE);
    fast arithmetic (FMA contraction allowed) within the fp32 tolerance of the reference-order fields."""
    entry = load_norms()[name]
    cfg = entry["config"]
    a = make_grid(cfg, flags=abi.ARITH_REFERENCE | abi.HOST_MIRROR_FULL | abi.L2_REFERENCE)
    a.run(library=cuda_lib)
    norms = a.convergence()
    assert ["%.10f" % norms["%s_l2" % f] for f in cfg["fields"]] == entry["l2_printed"]
    fa = fields_of(a)
    a.free()
    b = make_grid(cfg, flags=abi.ARITH_FAST | abi.HOST_MIRROR_FULL)
    b.run(library=cuda_lib)
    fb = fields_of(b)
    b.free()
    for k in range(fa.shape[0]):
        err = np.sqrt(((fb[k].astype(np.float64) - fa[k]) ** 2).sum() / (fa[k].astype(np.float64) ** 2).sum())
        assert err <= 1e-5


@pytest.mark.gpu
def test_generic_compile_error_is_reported(cuda_lib):
    import ctypes
    g = make_grid(KERNELS["gen_small_damped_so4_f32"]["config"], flags=abi.ARITH_REFERENCE | abi.HOST_MIRROR_FULL)
    p, keep = g.build_params()
    bad = g.generic_source.replace("opesci_generic_step", "opesci_generic_step_x").encode()
    p.generic_source = bad
    assert cuda_lib.opesci_b200_configure(ctypes.byref(p)) == 0
    grid = abi.OpesciGrid()
    assert cuda_lib.opesci_execute(ctypes.byref(grid), None) != 0
    assert b"opesci_generic_step" in cuda_lib.opesci_b200_last_error()
