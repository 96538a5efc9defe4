"""Worker for tests/test_slab_gloo.py: one rank of an x-slab run of the CPU oracle (gloo backend).

TEST INFRASTRUCTURE.  Launched by torch.distributed.run; writes the rank's local slab of every field
(both time levels), its slab geometry and its L2 partial norms to <outdir>/rank<r>.npz.
argv[2] is one configuration (JSON object) or a JSON list of them: job i of a list writes to
<outdir>/job<i>/, so that one rendezvous serves a whole batch of cases.
"""
import ctypes
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from common import make_grid  # noqa: E402
from opesci_fd_b200 import abi  # noqa: E402

EXCHANGE_FN = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                               ctypes.c_void_p, ctypes.c_size_t)


def main():
    outdir, cfgs = sys.argv[1], json.loads(sys.argv[2])
    use_cuda = len(sys.argv) > 3 and sys.argv[3] == "cuda"
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    if use_cuda:
        # product path: CUDA library, halo exchange by NCCL inside the library
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", rank)))
        lib = abi.load_library()
        ident = torch.zeros(abi.COMM_ID_BYTES, dtype=torch.uint8)
        if rank == 0:
            buf = (ctypes.c_ubyte * abi.COMM_ID_BYTES)()
            assert lib.opesci_b200_comm_unique_id(buf, abi.COMM_ID_BYTES) == 0, lib.opesci_b200_last_error()
            ident = torch.tensor(list(buf), dtype=torch.uint8)
        dist.broadcast(ident, 0)
        buf = (ctypes.c_ubyte * abi.COMM_ID_BYTES)(*ident.tolist())
        assert lib.opesci_b200_comm_init(rank, world, buf, abi.COMM_ID_BYTES) == 0, lib.opesci_b200_last_error()
    else:
        lib = abi.bind(ctypes.CDLL(os.path.join(ROOT, "oracle", "libopesci_oracle.so")))

    def view(ptr, n):
        return torch.frombuffer((ctypes.c_ubyte * n).from_address(ptr), dtype=torch.uint8)

    def exchange(user, send_lo, recv_lo, send_hi, recv_hi, nbytes):
        reqs = []
        if send_lo:
            reqs.append(dist.isend(view(send_lo, nbytes), rank - 1))
            reqs.append(dist.irecv(view(recv_lo, nbytes), rank - 1))
        if send_hi:
            reqs.append(dist.isend(view(send_hi, nbytes), rank + 1))
            reqs.append(dist.irecv(view(recv_hi, nbytes), rank + 1))
        for r in reqs:
            r.wait()
        return 0
    cb = EXCHANGE_FN(exchange)
    if not use_cuda:
        lib.opesci_oracle_set_exchange.argtypes = [EXCHANGE_FN, ctypes.c_void_p]
        lib.opesci_oracle_set_exchange(cb, None)

    if isinstance(cfgs, dict):
        run_job(outdir, cfgs, lib, rank, world, use_cuda)
    else:
        for i, cfg in enumerate(cfgs):
            job_dir = os.path.join(outdir, "job%d" % i)
            os.makedirs(job_dir, exist_ok=True)
            run_job(job_dir, cfg, lib, rank, world, use_cuda)
            dist.barrier()
    if use_cuda:
        lib.opesci_b200_comm_finalize()
    dist.barrier()
    dist.destroy_process_group()


def run_job(outdir, cfg, lib, rank, world, use_cuda):
    flags = os.environ.get("OPESCI_TEST_FLAGS")
    grid = make_grid(cfg, flags=int(flags) if (flags and use_cuda) else None)
    if cfg["kind"] == "eigenwave3d_read":
        # hand over only the planes this rank stores (its slab + halos), as a large run would
        l0, l1 = ctypes.c_int(), ctypes.c_int()
        assert lib.opesci_b200_slab_range(rank, world, grid.dim[0].value, cfg["so"], ctypes.byref(l0), ctypes.byref(l1)) == 0
        rho, vp, vs = grid.media_arrays
        grid.set_media_arrays(rho[l0.value:l1.value], vp[l0.value:l1.value], vs[l0.value:l1.value], plane0=l0.value)
    if cfg.get("hooks"):
        grid.set_receivers(cfg["hooks"]["receivers"])
        grid.set_source(cfg["hooks"]["source"], np.array(cfg["hooks"]["wavelet"], dtype=np.float32))
    orig = grid.build_params

    def with_slab():
        p, keep = orig()
        p.slab_rank, p.slab_nranks = rank, world
        return p, keep
    grid.build_params = with_slab
    if cfg.get("vts_prefix") and use_cuda:
        lib.opesci_b200_set_output(cfg["vts_prefix"].encode(), 0, cfg.get("vts_every", 1))
    if cfg.get("split_refresh"):
        os.environ["OPESCI_ORACLE_SPLIT_REFRESH"] = "1"   # oracle: the CUDA slab loop's refresh order
    else:
        os.environ.pop("OPESCI_ORACLE_SPLIT_REFRESH", None)
    grid.run(library=lib)
    transport = lib.opesci_b200_halo_transport() if use_cuda else 0
    if cfg.get("vts_prefix") and use_cuda:
        lib.opesci_b200_set_output(None, 0, 0)
    p = grid._params
    # local slab geometry (same arithmetic as include/opesci_slab.h)
    m, gdim = p.so // 2, p.dim[0]
    need = m if p.kind == abi.KIND_REGULAR_ACOUSTIC else (2 * m + 3 if p.so == 4 else 2 * m)   # include/opesci_slab.h
    H = max(abi.SLAB_HALO, need)
    n_int = gdim - 2 * m
    base, rem = divmod(n_int, world)
    X0 = m + rank * base + min(rank, rem)
    X1 = X0 + base + (1 if rank < rem else 0)
    L0 = 0 if rank == 0 else X0 - H
    L1 = gdim if rank == world - 1 else X1 + H
    own_lo = 0 if rank == 0 else X0
    own_hi = gdim if rank == world - 1 else X1
    n = p.nlevels * (L1 - L0) * p.dim[1] * p.dim[2]
    ctype = ctypes.c_double if p.is_double else ctypes.c_float
    fields = []
    for k in range(p.nfields):
        buf = ctypes.cast(grid._arg_grid.field[k], ctypes.POINTER(ctype * n)).contents
        fields.append(np.frombuffer(buf, dtype=np.float64 if p.is_double else np.float32).reshape(
            p.nlevels, L1 - L0, p.dim[1], p.dim[2]).copy())
    l2 = np.array(grid.convergence_f64())
    rec = grid.receiver_data() if hasattr(grid, 'receiver_data') else None
    np.savez(os.path.join(outdir, "rank%d.npz" % rank), fields=np.stack(fields), L0=L0, L1=L1, own_lo=own_lo, own_hi=own_hi, l2=l2,
             receivers=rec if rec is not None else np.zeros(0), transport=transport)
    grid.free()


if __name__ == "__main__":
    main()
