"""Development aid: BASELINE config 5 -- heterogeneous eigenwave3d, (n*world) x n x n cells, so=4 fp32,
x-slabs over `world` GPUs (run under torch.distributed.run).  Prints whole-job Gpts/s (max over ranks)."""
import ctypes
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import eigenwave3d as drv  # noqa: E402
from opesci_fd_b200 import abi  # noqa: E402
from opesci_fd_b200.util import synthetic_media_planes  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    steps, warm = 30, 5
    world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    lib = abi.load_library()
    ident = torch.zeros(abi.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
    if rank == 0:
        buf = (ctypes.c_ubyte * abi.COMM_ID_BYTES)()
        assert lib.opesci_b200_comm_unique_id(buf, abi.COMM_ID_BYTES) == 0
        ident = torch.tensor(list(buf), dtype=torch.uint8, device="cuda")
    dist.broadcast(ident, 0)
    buf = (ctypes.c_ubyte * abi.COMM_ID_BYTES)(*ident.cpu().tolist())
    assert lib.opesci_b200_comm_init(rank, world, buf, abi.COMM_ID_BYTES) == 0, lib.opesci_b200_last_error()
    nx = n * world
    dt = 0.4 / n / 1.5
    g = drv.eigenwave3d((float(world), 1.0, 1.0), (nx, n, n), dt, dt * (steps + warm), accuracy_order=[2, 4, 4, 4],
                        o_converge=False, read=True, rho_file="-", vp_file="-", vs_file="-", verbose=False)
    g.ntsteps.value = steps + warm
    g.b200_flags = abi.ARITH_FAST | abi.HOST_MIRROR_NONE
    dims = [d.value for d in g.dim]
    l0, l1 = ctypes.c_int(), ctypes.c_int()
    assert lib.opesci_b200_slab_range(rank, world, dims[0], 4, ctypes.byref(l0), ctypes.byref(l1)) == 0
    t0 = time.time()
    g.set_media_arrays(*synthetic_media_planes(dims, l0.value, l1.value - l0.value), plane0=l0.value)
    tgen = time.time() - t0
    orig = g.build_params

    def with_slab():
        p, k = orig()
        p.warmup_steps, p.slab_rank, p.slab_nranks = warm, rank, world
        return p, k
    g.build_params = with_slab
    dist.barrier()
    g.run(library=lib)
    secs, pts, launches = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
    lib.opesci_b200_last_timing(ctypes.byref(secs), ctypes.byref(pts), ctypes.byref(launches))
    t = torch.tensor([secs.value], device="cuda", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    l2 = g.convergence_f64()
    if rank == 0:
        gpts = pts.value * steps / float(t.item()) / 1e9
        print("hetero slabs: %dx%dx%d cells on %d GPUs, so=4 fp32: %.2f Gpts/s (%.2f ms/step), %.0f GB/s algorithmic per GPU "
              "(104 B/pt); media generation %.0fs; finite norms: %s"
              % (nx, n, n, world, gpts, float(t.item()) / steps * 1e3, gpts * 104 / world, tgen,
                 all(v == v and v < 1.0 for v in l2)), flush=True)
    g.free()
    lib.opesci_b200_comm_finalize()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
