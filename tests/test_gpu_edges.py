"""GPU: edge cases of the hot path, CUDA (reference arithmetic, through the C ABI) against the oracle bit for bit.

The reference has no unit tests; these cover what its loops do at the corners of their domain:
zero and one time step (time-level parity `ti = ntsteps % 2`), grids barely larger than the stencil, sizes that are
not multiples of any tile, every kernel family (fused so<=4, TMA-tiled so>=6 / fp64, regular acoustic incl. the
marching kernel), and heterogeneous media on ragged sizes."""
import numpy as np
import pytest

from common import bits, fields_of, make_grid
from opesci_fd_b200 import abi

pytestmark = pytest.mark.gpu

CASES = [
    # kind, so, grid_size, steps, double
    ("eigenwave3d", 4, [30, 28, 34], 0, False),      # no step at all: init + initial BC pass only
    ("eigenwave3d", 4, [30, 28, 34], 1, False),      # one step: results live on level 1
    ("eigenwave3d", 4, [1, 1, 1], 3, False),         # smallest legal grid (dim = 6): 2 interior points per axis
    ("eigenwave3d", 2, [2, 1, 3], 4, False),
    ("eigenwave3d", 12, [3, 2, 1], 3, False),        # dim 16 x 15 x 14 with m = 6
    ("eigenwave3d", 2, [61, 45, 67], 6, False),      # so=2 through the fused kernel, ragged sizes
    ("eigenwave3d", 4, [129, 17, 63], 5, False),     # long in x, thin in y
    ("eigenwave3d", 4, [17, 131, 12], 5, False),     # z shorter than one tile
    ("eigenwave3d", 6, [37, 41, 43], 5, False),      # tiled kernels, odd m
    ("eigenwave3d", 10, [40, 33, 70], 4, True),      # tiled kernels, fp64, odd m
    ("eigenwave3d", 12, [45, 38, 36], 4, True),
    ("eigenwave3d", 4, [50, 44, 48], 7, True),       # so=4 fp64: tiled, Levander
    ("simplewave3d", 2, [33, 29, 31], 9, False),
    ("simplewave3d", 12, [40, 37, 35], 8, True),
    ("simplewave3d", 8, [2, 2, 2], 5, False),
    ("eigenwave3d_read", 4, [35, 67, 29], 9, False),  # heterogeneous, fused
    ("eigenwave3d_read", 2, [31, 30, 40], 6, False),
    ("eigenwave3d_read", 6, [33, 36, 34], 5, False),  # heterogeneous, tiled, Robertsson
    ("eigenwave3d_read", 4, [2, 3, 1], 3, False),     # heterogeneous on a minimal grid
    # z strip of the fused path: interior z extent = 2 tiles of 60 + a few columns done per point (opesci_b200.cu:fused)
    ("eigenwave3d", 4, [21, 26, 124], 6, False),      # 125 = 2 x 60 + 5 (the 1024^3 remainder)
    ("eigenwave3d", 4, [20, 23, 120], 5, False),      # 121 = 2 x 60 + 1
    ("eigenwave3d", 4, [19, 22, 127], 5, False),      # 128 = 2 x 60 + 8 (widest strip)
    ("eigenwave3d", 2, [18, 21, 123], 5, False),      # so=2: tiles of 60, 124 = 2 x 60 + 4
    ("eigenwave3d_read", 4, [17, 20, 125], 6, False), # heterogeneous strip
]


@pytest.mark.parametrize("kind,so,size,steps,double", CASES)
def test_cuda_equals_oracle_on_edge_cases(kind, so, size, steps, double, cuda_lib, oracle_lib):
    cfg = dict(kind=kind, so=so, grid_size=size, dt=0.0015, steps=steps, double=double,
               domain=[1.0, 0.9, 1.2], rho=1.1, vp=1.9, vs=1.0, seed=3)
    a, b = make_grid(cfg, flags=abi.ARITH_REFERENCE | abi.HOST_MIRROR_FULL), make_grid(cfg)
    a.run(library=cuda_lib)
    b.run(library=oracle_lib)
    fa, fb = fields_of(a), fields_of(b)
    assert fa.shape == fb.shape
    assert not np.isnan(fb).any()
    assert int((bits(fa) != bits(fb)).sum()) == 0
    np.testing.assert_allclose(a.convergence_f64(), b.convergence_f64(), rtol=1e-12, atol=0)
    a.free()
    b.free()


def test_fast_arithmetic_on_device_pointers(cuda_lib):
    """HOST_MIRROR_NONE: grid->field[] are device pointers; opesci_convergence reduces on the device and equals the
    host-mirror run of the same model."""
    cfg = dict(kind="eigenwave3d", so=4, grid_size=[48, 40, 52], dt=0.002, steps=10, double=False, domain=[1.0, 1.0, 1.0])
    a = make_grid(cfg, flags=abi.ARITH_FAST | abi.HOST_MIRROR_NONE)
    b = make_grid(cfg, flags=abi.ARITH_FAST | abi.HOST_MIRROR_FULL)
    a.run(library=cuda_lib)
    b.run(library=cuda_lib)
    assert a.convergence_f64() == b.convergence_f64()
    a.free()
    b.free()


FACE_SUBSETS = [
    # so, size, faces (dimension 1..3, side 0/1)
    (4, [40, 36, 70], [(2, 0), (2, 1), (3, 0), (3, 1)]),      # no x faces; both z faces: fused kernel with the z-fold
    (4, [40, 36, 70], [(1, 0), (3, 0)]),                      # a single z face: z-fold off, per-face ghost kernels
    (4, [33, 30, 64], [(1, 1), (2, 0), (3, 1)]),
    (4, [30, 28, 34], []),                                    # no free surface at all: interior updates only
    (8, [37, 41, 43], [(1, 0), (2, 1), (3, 0)]),              # Robertsson, tiled kernels
    (4, [35, 31, 39], [(2, 1), (3, 0), (3, 1)]),              # fp64 below
]


@pytest.mark.parametrize("so,size,faces", FACE_SUBSETS)
@pytest.mark.parametrize("double", [False, True])
def test_free_surface_on_a_subset_of_the_faces(so, size, faces, double, cuda_lib, oracle_lib):
    """set_free_surface_boundary on some faces only (reference: opesci/staggeredgrid.py:214-232; loops of faces without
    boundary code are not emitted, :766-768).  PARITY UNPINNED by the reference: its generator crashes for a subset
    (field.bc stays None, fields.py:36 -> TypeError in transform_bc, staggeredgrid.py:176), so this is CUDA == oracle."""
    cfg = dict(kind="eigenwave3d", so=so, grid_size=size, dt=0.0015, steps=7, double=double,
               domain=[1.0, 0.9, 1.2], rho=1.1, vp=1.9, vs=1.0, faces=faces)
    a, b = make_grid(cfg, flags=abi.ARITH_REFERENCE | abi.HOST_MIRROR_FULL), make_grid(cfg)
    a.run(library=cuda_lib)
    b.run(library=oracle_lib)
    fa, fb = fields_of(a), fields_of(b)
    assert not np.isnan(fb).any()
    assert int((bits(fa) != bits(fb)).sum()) == 0
    # and the boundary treatment really is off on the other faces: it differs from the all-faces run
    if len(faces) < 6:
        c = make_grid(dict(cfg, faces=[(d, s) for d in (1, 2, 3) for s in (0, 1)]))
        c.run(library=oracle_lib)
        assert int((bits(fields_of(c)) != bits(fb)).sum()) > 0
        c.free()
    a.free()
    b.free()


@pytest.mark.parametrize("arith", [abi.ARITH_REFERENCE, abi.ARITH_FAST])
@pytest.mark.parametrize("kind", ["eigenwave3d", "eigenwave3d_read"])
def test_fused_path_variants_are_bit_identical(kind, arith, cuda_lib):
    """The shipped schedule (z-face loops in the z-edge tiles, interior tiles as y-stacked 2-CTA clusters exchanging their
    inner halo rows through distributed shared memory) against the same kernel without clusters (NO_PAIR) and with the
    z-face loops back in the separate face kernels (NO_ZFOLD): every cell of every field on both levels identical, in both
    arithmetic modes -- the variants change where bytes travel, never an operand or an operation."""
    cfg = dict(kind=kind, so=4, grid_size=[70, 95, 190], dt=0.0015, steps=11, double=False,
               domain=[1.0, 0.9, 1.2], rho=1.1, vp=1.9, vs=1.0, seed=11)
    out = []
    for extra in (0, abi.NO_PAIR, abi.NO_PAIR | abi.NO_ZFOLD):
        g = make_grid(cfg, flags=arith | abi.HOST_MIRROR_FULL | extra)
        g.run(library=cuda_lib)
        out.append(fields_of(g))
        g.free()
    assert int((bits(out[0]) != bits(out[1])).sum()) == 0
    assert int((bits(out[0]) != bits(out[2])).sum()) == 0
