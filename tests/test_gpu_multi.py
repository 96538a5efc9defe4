"""GPU, 2 ranks, one process per GPU: the slab-decomposed CUDA run equals the single-GPU run bit for bit -- with the
default halo transport (the neighbour's fields mapped with cudaIpc, planes pulled by the copy engines, NCCL tokens for
the ordering) and with the planes sent by ncclSend/ncclRecv (OPESCI_HALO_P2P=0)."""
import json
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from common import ROOT, bits, fields_of, make_grid
from opesci_fd_b200 import abi

pytestmark = pytest.mark.gpu


def _ngpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.skipif(_ngpus() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
@pytest.mark.parametrize("arith,kind,so", [(abi.ARITH_REFERENCE, "eigenwave3d", 4), (abi.ARITH_FAST, "eigenwave3d", 4),
                                           (abi.ARITH_REFERENCE, "eigenwave3d_read", 4), (abi.ARITH_REFERENCE, "simplewave3d", 4),
                                           (abi.ARITH_REFERENCE, "eigenwave3d", 8), (abi.ARITH_REFERENCE, "eigenwave3d", 12)])
def test_two_gpu_slabs_equal_single_gpu(arith, kind, so, cuda_lib, tmp_path):
    _slabs_equal_single(arith, kind, so, [96, 70, 130], cuda_lib, tmp_path)


@pytest.mark.skipif(_ngpus() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
@pytest.mark.parametrize("kind", ["eigenwave3d", "eigenwave3d_read"])
def test_two_gpu_slabs_with_z_strip(kind, cuda_lib, tmp_path):
    """interior z extent 125 = 2 tiles + 5 columns: the per-point z strip of the fused path, chunk by chunk"""
    _slabs_equal_single(abi.ARITH_REFERENCE, kind, 4, [96, 40, 124], cuda_lib, tmp_path)


@pytest.mark.skipif(_ngpus() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
@pytest.mark.parametrize("kind,so", [("eigenwave3d", 4), ("eigenwave3d_read", 4), ("simplewave3d", 4), ("eigenwave3d", 8)])
def test_two_gpu_slabs_planes_by_nccl_send_recv(kind, so, cuda_lib, tmp_path, monkeypatch):
    monkeypatch.setenv("OPESCI_HALO_P2P", "0")
    _slabs_equal_single(abi.ARITH_REFERENCE, kind, so, [96, 70, 130], cuda_lib, tmp_path, transport=1)


def _slabs_equal_single(arith, kind, so, size, cuda_lib, tmp_path, transport=2):
    cfg = dict(kind=kind, so=so, grid_size=size, dt=0.002, steps=9, double=False,
               domain=[1.0, 0.9, 0.8], rho=1.2, vp=1.6, vs=0.8, seed=5)
    single = make_grid(cfg, flags=arith | abi.HOST_MIRROR_FULL)
    single.run(library=cuda_lib)
    ref = fields_of(single)
    ref_l2 = np.array(single.convergence_f64())
    single.free()
    os.environ["OPESCI_TEST_FLAGS"] = str(arith | abi.HOST_MIRROR_FULL)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "slab_worker.py"), str(tmp_path), json.dumps(cfg), "cuda"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    covered = 0
    for r in range(2):
        z = np.load(os.path.join(str(tmp_path), "rank%d.npz" % r))
        L0, own_lo, own_hi = int(z["L0"]), int(z["own_lo"]), int(z["own_hi"])
        mine = np.ascontiguousarray(z["fields"][:, :, own_lo - L0:own_hi - L0])
        want = np.ascontiguousarray(ref[:, :, own_lo:own_hi])
        assert int((bits(mine) != bits(want)).sum()) == 0, "rank %d differs from the single-GPU run" % r
        covered += own_hi - own_lo
        assert int(z["transport"]) == transport, "halo transport %d, expected %d (include/opesci_b200.h)" % (int(z["transport"]), transport)
        # every rank holds the all-reduced global norms
        np.testing.assert_allclose(z["l2"], ref_l2, rtol=1e-12)
    assert covered == ref.shape[2]


@pytest.mark.skipif(_ngpus() < 2, reason="needs 2 GPUs (run with gpurun --gpus 2)")
@pytest.mark.parametrize("so", [4, 8])
def test_two_gpu_vts_pieces_tile_the_single_gpu_snapshot(so, cuda_lib, tmp_path):
    """per-step field output with slabs: every rank writes the planes it owns; together they are the single-GPU file"""
    from test_io import read_vts
    cfg = dict(kind="eigenwave3d", so=so, grid_size=[96, 30, 34], dt=0.002, steps=4, double=False, domain=[1.0, 0.9, 0.8])
    one = str(tmp_path / "one_")
    cuda_lib.opesci_b200_set_output(one.encode(), 0, 2)
    single = make_grid(cfg, flags=abi.ARITH_REFERENCE | abi.HOST_MIRROR_NONE)
    single.run(library=cuda_lib)
    single.free()
    cuda_lib.opesci_b200_set_output(None, 0, 0)
    os.environ["OPESCI_TEST_FLAGS"] = str(abi.ARITH_REFERENCE | abi.HOST_MIRROR_FULL)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(ROOT, "tests", "slab_worker.py"), str(tmp_path), json.dumps(dict(cfg, vts_prefix=str(tmp_path / "two_"), vts_every=2)), "cuda"]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    for ti in (0, 2):
        _, f1, p1 = read_vts("%s%d.vts" % (one, ti))
        parts = [read_vts("%s%d_r%d.vts" % (str(tmp_path / "two_"), ti, r)) for r in range(2)]
        f2 = np.concatenate([p[1] for p in parts])
        p2 = np.concatenate([p[2] for p in parts])
        assert np.array_equal(f1.view(np.uint32), f2.view(np.uint32)) and np.array_equal(p1, p2)
