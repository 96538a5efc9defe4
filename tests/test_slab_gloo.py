"""CPU, world_size 2 and 3 over gloo: the x-slab decomposition (include/opesci_slab.h) is exact.

Every rank runs the oracle on its slab (+8 halo planes per inner side, refreshed once per step through a
gloo send/recv callback); the owned planes of all ranks, stitched together, must be BIT-IDENTICAL to the
single-domain run, and the per-slab L2 sums must add up to the global norms.  The CUDA library uses the
same geometry and loop-range rules with NCCL in place of the callback (tests/test_gpu_multi.py)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from common import ROOT, bits, fields_of, make_grid


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


CASES = [(2, 4, "eigenwave3d"), (3, 4, "eigenwave3d"), (2, 2, "eigenwave3d"),
         (2, 4, "eigenwave3d_read"), (2, 4, "simplewave3d"), (3, 8, "simplewave3d"),
         (2, 8, "eigenwave3d"), (2, 12, "eigenwave3d"), (2, 6, "eigenwave3d_read")]
# the same decomposition with the refresh order of the CUDA slab loop: stress planes after the stress ghost loops,
# velocity planes after the velocity ghost loops (oracle: OPESCI_ORACLE_SPLIT_REFRESH)
SPLIT_CASES = [(2, 4, "eigenwave3d"), (3, 4, "eigenwave3d"), (2, 2, "eigenwave3d"), (2, 4, "eigenwave3d_read"), (3, 8, "eigenwave3d")]
HOOKS = dict(receivers=[[0.3, 0.4, 0.4], [1.2, 0.5, 0.3], [1.9, 0.2, 0.6], [1.0, 0.45, 0.4]], source=[1.0, 0.45, 0.4],
             wavelet=[0.0, 0.01, 0.03, 0.01, -0.02, 0.0])


def _case_cfg(world, so, kind):
    return dict(kind=kind, so=so, grid_size=[30 * world, 11, 9], dt=0.002, steps=11, double=False,
                domain=[1.0 * world, 0.9, 0.8], rho=1.2, vp=1.6, vs=0.8, seed=11)


def _hooks_cfg():
    return dict(kind="eigenwave3d", so=4, grid_size=[60, 11, 9], dt=0.002, steps=11, double=False,
                domain=[2.0, 0.9, 0.8], rho=1.2, vp=1.6, vs=0.8, hooks=HOOKS)


@pytest.fixture(scope="module")
def slab_runs(oracle_lib, tmp_path_factory):
    """One torch.distributed.run per world size serves every case of that size (a rendezvous plus two or three
    `import torch` cost more than all the oracle runs together).  -> {case key: directory with rank<r>.npz}"""
    jobs = {}
    for world, so, kind in CASES:
        jobs.setdefault(world, []).append(((world, so, kind), _case_cfg(world, so, kind)))
    for world, so, kind in SPLIT_CASES:
        jobs.setdefault(world, []).append((("split", world, so, kind), dict(_case_cfg(world, so, kind), split_refresh=True)))
    jobs[2].append(("hooks", _hooks_cfg()))
    where = {}
    for world, batch in sorted(jobs.items()):
        out_root = tmp_path_factory.mktemp("slabs_w%d" % world)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world),
               "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
               os.path.join(ROOT, "tests", "slab_worker.py"), str(out_root), json.dumps([cfg for _, cfg in batch])]
        out = subprocess.run(cmd, env=dict(os.environ, OMP_NUM_THREADS="2"), capture_output=True, text=True, timeout=900)
        assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
        for i, (key, _) in enumerate(batch):
            where[key] = os.path.join(str(out_root), "job%d" % i)
    return where


@pytest.mark.parametrize("world,so,kind", CASES)
def test_slab_decomposition_is_bit_exact(world, so, kind, oracle_lib, slab_runs):
    _check_against_single(world, so, kind, oracle_lib, slab_runs[(world, so, kind)])


@pytest.mark.parametrize("world,so,kind", SPLIT_CASES)
def test_slab_decomposition_with_split_refresh_is_bit_exact(world, so, kind, oracle_lib, slab_runs):
    """The CUDA slab loop refreshes the stress halo planes while the velocity shell and the velocity ghost loops run: a
    kernel may read such a plane before or after it changes.  "Before" is the default schedule above, "after" is this
    one; owned planes are bit-identical either way, so the overlap on the GPU is exact."""
    _check_against_single(world, so, kind, oracle_lib, slab_runs[("split", world, so, kind)])


def _check_against_single(world, so, kind, oracle_lib, outdir):
    single = make_grid(_case_cfg(world, so, kind))
    single.run(library=oracle_lib)
    ref = fields_of(single)
    ref_l2 = np.array(single.convergence_f64())
    single.free()
    sums = np.zeros(ref.shape[0])
    covered = 0
    for r in range(world):
        z = np.load(os.path.join(outdir, "rank%d.npz" % r))
        L0, own_lo, own_hi = int(z["L0"]), int(z["own_lo"]), int(z["own_hi"])
        mine = z["fields"][:, :, own_lo - L0:own_hi - L0]
        want = ref[:, :, own_lo:own_hi]
        assert int((bits(np.ascontiguousarray(mine)) != bits(np.ascontiguousarray(want))).sum()) == 0, "rank %d" % r
        covered += own_hi - own_lo
        sums += z["l2"] ** 2
    assert covered == ref.shape[2]
    np.testing.assert_allclose(np.sqrt(sums), ref_l2, rtol=1e-12)


def test_slab_point_source_and_receivers(oracle_lib, slab_runs):
    """Source and receivers with slabs: every rank handles the cells on planes it owns, so the per-rank receiver
    traces add up to the single-domain traces and the fields stay bit-identical."""
    single = make_grid(_hooks_cfg())
    single.set_receivers(HOOKS["receivers"])
    single.set_source(HOOKS["source"], np.array(HOOKS["wavelet"], dtype=np.float32))
    single.run(library=oracle_lib)
    ref = fields_of(single)
    ref_rec = single.receiver_data().copy()
    single.free()
    outdir = slab_runs["hooks"]
    total = np.zeros_like(ref_rec)
    for r in range(2):
        z = np.load(os.path.join(outdir, "rank%d.npz" % r))
        L0, own_lo, own_hi = int(z["L0"]), int(z["own_lo"]), int(z["own_hi"])
        mine = np.ascontiguousarray(z["fields"][:, :, own_lo - L0:own_hi - L0])
        assert int((bits(mine) != bits(np.ascontiguousarray(ref[:, :, own_lo:own_hi]))).sum()) == 0, "rank %d" % r
        total += z["receivers"]
    assert int((bits(total) != bits(ref_rec)).sum()) == 0
    assert np.abs(ref_rec).max() > 0
