"""GPU (one device is enough): slab exactness of the CUDA path, proved in loopback.

`opesci_b200_execute_loopback` runs N logical ranks concurrently on one GPU, each on its own host thread through the
same schedule a real NCCL rank runs (x-chunk table, halo exchange on its own stream overlapped with the middle chunks,
end chunks behind `ev_join`, ghost loops, shell); only the transport of the halo planes differs (device-to-device copies
instead of ncclSend / ncclRecv).  Every owned plane of every field on every time level must equal the single-domain run
BIT FOR BIT -- the dependency argument of SURVEY.md 8e / include/opesci_slab.h (reference loops: opesci/fields.py:208-242,
opesci/staggeredgrid.py:815-864).  The same comparison over real NCCL is tests/test_gpu_multi.py (needs 2 GPUs).
"""
import ctypes

import numpy as np
import pytest

from common import bits, fields_of, make_grid
from opesci_fd_b200 import abi

pytestmark = pytest.mark.gpu


def run_loopback(cfg, flags, nranks, lib):
    """-> list of (L0, L1, own_lo, own_hi, fields[nf][nlevels][L1-L0][dim2][dim3]) per rank"""
    g = make_grid(cfg, flags=flags | abi.HOST_MIRROR_FULL)
    params, keep = g.build_params()
    params.flags = int(g.b200_flags)
    assert lib.opesci_b200_configure(ctypes.byref(params)) == 0, lib.opesci_b200_last_error()
    grids = (abi.OpesciGrid * nranks)()
    rc = lib.opesci_b200_execute_loopback(nranks, grids)
    assert rc == 0, lib.opesci_b200_last_error().decode()
    dims = [params.dim[0], params.dim[1], params.dim[2]]
    m = params.so // 2
    nint = dims[0] - 2 * m
    acoustic = params.kind == abi.KIND_REGULAR_ACOUSTIC
    need = m if acoustic else (2 * m + 3 if params.so == 4 else 2 * m)
    halo = max(abi.SLAB_HALO, need)
    dtype, ctype = (np.float64, ctypes.c_double) if params.is_double else (np.float32, ctypes.c_float)
    out = []
    for r in range(nranks):
        # include/opesci_slab.h: opesci_slab_make
        base, rem = nint // nranks, nint % nranks
        X0 = m + r * base + min(r, rem)
        X1 = X0 + base + (1 if r < rem else 0)
        L0 = 0 if r == 0 else X0 - halo
        L1 = dims[0] if r == nranks - 1 else X1 + halo
        own_lo = 0 if r == 0 else X0
        own_hi = dims[0] if r == nranks - 1 else X1
        if not acoustic:
            l0, l1 = ctypes.c_int(), ctypes.c_int()
            assert lib.opesci_b200_slab_range(r, nranks, dims[0], params.so, ctypes.byref(l0), ctypes.byref(l1)) == 0
            assert (l0.value, l1.value) == (L0, L1)
        n = params.nlevels * (L1 - L0) * dims[1] * dims[2]
        fields = []
        for k in range(params.nfields):
            buf = ctypes.cast(grids[r].field[k], ctypes.POINTER(ctype * n)).contents
            fields.append(np.frombuffer(buf, dtype=dtype).reshape(params.nlevels, L1 - L0, dims[1], dims[2]).copy())
        out.append((L0, L1, own_lo, own_hi, np.stack(fields)))
        one = abi.OpesciGrid()
        ctypes.memmove(ctypes.byref(one), ctypes.byref(grids[r]), ctypes.sizeof(abi.OpesciGrid))
        assert lib.opesci_free(ctypes.byref(one)) == 0
    del keep
    return out


def check_against_single(cfg, flags, nranks, lib):
    single = make_grid(cfg, flags=flags | abi.HOST_MIRROR_FULL)
    single.run(library=lib)
    ref = fields_of(single)
    single.free()
    covered = 0
    for r, (L0, L1, own_lo, own_hi, mine) in enumerate(run_loopback(cfg, flags, nranks, lib)):
        a = np.ascontiguousarray(mine[:, :, own_lo - L0:own_hi - L0])
        b = np.ascontiguousarray(ref[:, :, own_lo:own_hi])
        nbad = int((bits(a) != bits(b)).sum())
        assert nbad == 0, "rank %d of %d: %d cells of its owned planes differ from the single-domain run" % (r, nranks, nbad)
        covered += own_hi - own_lo
    assert covered == ref.shape[2]


CASES = [
    # (id, kind, so, double, arith, size, nranks)
    ("so4_ref_2", "eigenwave3d", 4, False, abi.ARITH_REFERENCE, [96, 70, 130], 2),
    ("so4_fast_2", "eigenwave3d", 4, False, abi.ARITH_FAST, [96, 70, 130], 2),
    ("so4_ref_3", "eigenwave3d", 4, False, abi.ARITH_REFERENCE, [150, 40, 70], 3),
    ("so4_ref_4", "eigenwave3d", 4, False, abi.ARITH_REFERENCE, [200, 30, 66], 4),
    ("so4_zstrip", "eigenwave3d", 4, False, abi.ARITH_REFERENCE, [96, 40, 124], 2),
    ("hetero_2", "eigenwave3d_read", 4, False, abi.ARITH_REFERENCE, [96, 70, 130], 2),
    ("hetero_zstrip_3", "eigenwave3d_read", 4, False, abi.ARITH_REFERENCE, [150, 40, 124], 3),
    ("hetero_so8_2", "eigenwave3d_read", 8, False, abi.ARITH_REFERENCE, [96, 40, 70], 2),
    ("so8_2", "eigenwave3d", 8, False, abi.ARITH_REFERENCE, [96, 70, 130], 2),
    ("so12_2", "eigenwave3d", 12, False, abi.ARITH_REFERENCE, [96, 50, 70], 2),
    ("so2_2", "eigenwave3d", 2, False, abi.ARITH_REFERENCE, [96, 50, 70], 2),
    ("so4_f64_2", "eigenwave3d", 4, True, abi.ARITH_REFERENCE, [96, 50, 70], 2),
    ("so8_f64_3", "eigenwave3d", 8, True, abi.ARITH_FAST, [150, 40, 50], 3),
    ("acoustic_2", "simplewave3d", 4, False, abi.ARITH_REFERENCE, [96, 70, 130], 2),
    ("acoustic_so8_4", "simplewave3d", 8, False, abi.ARITH_REFERENCE, [128, 40, 66], 4),
]


@pytest.mark.parametrize("name,kind,so,double,arith,size,nranks", CASES, ids=[c[0] for c in CASES])
def test_loopback_slabs_equal_the_single_domain_run(name, kind, so, double, arith, size, nranks, cuda_lib):
    cfg = dict(kind=kind, so=so, grid_size=size, dt=0.002, steps=9, double=double,
               domain=[1.0, 0.9, 0.8], rho=1.2, vp=1.6, vs=0.8, seed=5)
    check_against_single(cfg, arith, nranks, cuda_lib)


def test_loopback_even_step_count_and_one_step(cuda_lib):
    for steps in (1, 2, 10):
        cfg = dict(kind="eigenwave3d", so=4, grid_size=[96, 40, 66], dt=0.002, steps=steps, double=False,
                   domain=[1.0, 0.9, 0.8], rho=1.2, vp=1.6, vs=0.8)
        check_against_single(cfg, abi.ARITH_REFERENCE, 2, cuda_lib)


def test_loopback_rejects_slabs_thinner_than_the_halo(cuda_lib):
    cfg = dict(kind="eigenwave3d", so=4, grid_size=[40, 20, 20], dt=0.002, steps=2, double=False, domain=[1.0, 1.0, 1.0])
    g = make_grid(cfg, flags=abi.ARITH_REFERENCE | abi.HOST_MIRROR_FULL)
    params, keep = g.build_params()
    assert cuda_lib.opesci_b200_configure(ctypes.byref(params)) == 0
    grids = (abi.OpesciGrid * 8)()
    assert cuda_lib.opesci_b200_execute_loopback(8, grids) != 0
    assert b"thinner than the halo" in cuda_lib.opesci_b200_last_error()
