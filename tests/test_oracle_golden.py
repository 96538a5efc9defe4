"""CPU: the oracle (C restatement) against the committed golden fixtures.

The fixtures were produced by the reference's OWN generated OpenMP C++ (tests/golden/make_golden.py,
oracle/refgen/make_ref.py).  The bar is bit-exactness of every cell of every field on both time
levels, and equality of the L2 norms to all printed digits -- for so = 2..12, fp32 and fp64,
staggered (Levander at so=4, Robertsson otherwise) and regular grids.
"""
import numpy as np
import pytest

import hashlib
import json
import os

from common import GOLDEN, bits, fields_of, golden_names, load_golden, make_grid

HASHES = json.load(open(os.path.join(GOLDEN, "hashes.json")))


# (gen_*: PDE systems compiled at run time -- pinned on the reference's own output directly, tests/test_generic.py)
@pytest.mark.parametrize("name", [n for n in golden_names() if not n.startswith("gen_")])
def test_oracle_bit_exact_vs_reference_golden(name, oracle_lib):
    cfg, ref_fields, ref_l2 = load_golden(name)
    grid = make_grid(cfg)
    grid.run(library=oracle_lib)
    mine = fields_of(grid)
    assert mine.shape == ref_fields.shape
    assert mine.dtype == ref_fields.dtype
    for k, fname in enumerate(cfg["fields"]):
        nbad = int((bits(mine[k]) != bits(ref_fields[k])).sum())
        assert nbad == 0, "%s: %d cells differ from the reference's generated code" % (fname, nbad)
    if cfg["kind"] == "eigenwave3d_read":
        grid.free()   # `read` mode runs with converge=False: the reference prints no norms
        return
    # reference-faithful norm arithmetic (serial accumulation in real_t): same printed digits
    norms = grid.convergence()
    got = np.array([norms["%s_l2" % f] for f in cfg["fields"]])
    assert ["%.9e" % v for v in got] == ["%.9e" % v for v in ref_l2]
    # the double-accumulated norms (what the CUDA library reports) agree to accumulation error
    got64 = np.array(grid.convergence_f64())
    np.testing.assert_allclose(got64, ref_l2, rtol=2e-5 if not cfg["double"] else 2e-9)  # golden norms carry 10 digits
    grid.free()


# (the 512^3 and so=8 256^3 entries are GPU-side fixtures: minutes on the CPU oracle; the 256^3 x 40-step one stays, ~1 min)
@pytest.mark.parametrize("name", sorted(n for n in HASHES if HASHES[n]["config"]["kind"] == "eigenwave3d_read" or n == "ew_large_so4_f32_n256"))
def test_oracle_bit_exact_vs_patched_reference_hashes(name, oracle_lib):
    """Heterogeneous `read` mode at 48x40x44 cells x 40 steps (random rho/vp/vs per cell): the fields are too
    large to commit, so the fixture holds the sha256 of the raw bits of every field as produced by the
    patched reference (oracle/refgen/make_ref.py)."""
    entry = HASHES[name]
    grid = make_grid(entry["config"])
    grid.run(library=oracle_lib)
    mine = fields_of(grid)
    assert not np.isnan(mine).any()
    got = [hashlib.sha256(mine[k].tobytes()).hexdigest() for k in range(mine.shape[0])]
    assert got == entry["sha256"]
    grid.free()
