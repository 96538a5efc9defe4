"""Point source + receivers: semantics of the reference's hand-written propagator
(tests/src/test_ref_iso_elastic.cpp:227-290) added at the end of every time step.

The generated code has no such hooks and the hand-written propagator cannot be run (its data files are absent), so
parity here is CUDA == oracle bit for bit plus properties: receivers without a source read exactly the field
samples of the unhooked run; a source changes the field linearly from its cell outwards."""
import numpy as np
import pytest

from common import bits, fields_of, make_grid
from opesci_fd_b200 import abi

CFG = dict(kind="eigenwave3d", so=4, grid_size=[44, 40, 36], dt=0.002, steps=12, double=False, domain=[1.0, 0.9, 0.8],
           rho=1.2, vp=1.6, vs=0.8)
RECEIVERS = [(0.5, 0.45, 0.4), (0.1, 0.1, 0.1), (0.93, 0.3, 0.72), (0.5, 0.45, 0.42)]


def _wavelet(n):
    t = np.arange(n, dtype=np.float32)
    return (np.exp(-((t - 4.0) / 2.0) ** 2) * 1e-2).astype(np.float32)


def _run(lib, cfg=CFG, source=True, receivers=True, flags=None):
    g = make_grid(cfg, flags=flags)
    if receivers:
        g.set_receivers(RECEIVERS)
    if source:
        g.set_source((0.5, 0.45, 0.4), _wavelet(8))
    g.run(library=lib)
    f = fields_of(g)
    rec = None if g.receiver_data() is None else g.receiver_data().copy()
    g.free()
    return f, rec


def test_oracle_receivers_sample_the_fields_and_source_is_local(oracle_lib):
    f0, _ = _run(oracle_lib, source=False, receivers=False)
    f1, rec1 = _run(oracle_lib, source=False, receivers=True)
    assert int((bits(f0) != bits(f1)).sum()) == 0              # receivers do not disturb the run
    g = make_grid(CFG)
    cells = [g._cell_of(c) for c in RECEIVERS]
    ti = CFG["steps"] - 1
    lvl = CFG["steps"] % 2                                     # the level written by the last step
    for r, (x, y, z) in enumerate(cells):
        assert rec1[ti, 0, r] == f1[0, lvl, x, y, z]
        assert rec1[ti, 3, r] == np.float32((f1[3, lvl, x, y, z] + f1[4, lvl, x, y, z] + f1[5, lvl, x, y, z]) / np.float32(3))
    f2, rec2 = _run(oracle_lib, source=True, receivers=True)
    assert np.abs(rec2[:, :, 0] - rec1[:, :, 0]).max() > 0      # the receiver at the source hears it
    # locality: after 3 steps the disturbance has travelled at most 2m cells per step from the source cell
    short = dict(CFG, steps=3)
    a, _ = _run(oracle_lib, short, source=False, receivers=False)
    b, _ = _run(oracle_lib, short, source=True, receivers=False)
    diff = np.abs(b.astype(np.float64) - a).max(axis=(0, 1))    # [x][y][z]
    sx, sy, sz = cells[0]
    assert diff[sx, sy, sz] > 0
    reach = 2 * 2 * 3
    far = np.ones_like(diff, dtype=bool)
    far[sx - reach:sx + reach + 1, sy - reach:sy + reach + 1, sz - reach:sz + reach + 1] = False
    assert far.any() and diff[far].max() == 0


@pytest.mark.gpu
@pytest.mark.parametrize("kind,so,double", [("eigenwave3d", 4, False), ("eigenwave3d", 8, False), ("eigenwave3d", 4, True),
                                           ("eigenwave3d_read", 4, False)])
def test_cuda_hooks_equal_oracle(kind, so, double, cuda_lib, oracle_lib):
    cfg = dict(CFG, kind=kind, so=so, double=double, seed=9)
    fo, ro = _run(oracle_lib, cfg)
    fc, rc = _run(cuda_lib, cfg, flags=abi.ARITH_REFERENCE | abi.HOST_MIRROR_FULL)
    assert int((bits(fo) != bits(fc)).sum()) == 0
    assert ro.shape == rc.shape == (cfg["steps"], 4, len(RECEIVERS))
    assert int((bits(ro) != bits(rc)).sum()) == 0
