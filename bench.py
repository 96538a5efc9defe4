#!/usr/bin/env python
"""bench.py -- grid-point updates per second of the elastic eigenwave3d time loop on B200.

Metric (BASELINE.json / SURVEY.md 8d): Gpts/s = (N1+1)(N2+1)(N3+1) * steps / t_loop / 1e9 for
eigenwave3d, so=4, fp32; one grid-point update = all 9 fields advanced one leapfrog step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--n 1024] [--arith fast|reference]
  python bench.py --impl reference ...      the reference's own generated OpenMP C++ on the host cores

One "step" is one leapfrog time step over the whole grid.  `value` is measured with the fields
resident in HBM (CUDA events around exactly K steps after W warm-up steps, inside the library, on
the launching stream; the 78 GB working set is far larger than L2).  `e2e` is the same metric
through the reference-facing C ABI call `opesci_execute` with HOST result arrays: allocation,
initialisation, W+K steps and the device->host copy of all 18 level arrays are inside its timed
region.  `roofline` is computed for the dominant kernel from SURVEY.md 8d's 72 B per point update.
`cpu_baseline` times the reference's generated code (oracle/_ref, g++ -O3 -fopenmp) on this box's
host cores on a bounded sample (512^3, time loop isolated by differencing two step counts).
"""
import argparse
import ctypes
import json
import os
import re
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

METRIC = "Gpts/s elastic eigenwave3d so=4 fp32"
BYTES_PER_POINT = 72.0          # 9 fields x (1 read + 1 write) x 4 B


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.check_output(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                               "--format=csv,noheader,nounits"], timeout=5).decode().strip()
                self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        busy = [v for v in sm if v > 0]
        return {"sm_mhz": statistics.median(busy) if busy else None,
                "sm_max_mhz": float(self.rows[0][1]) if self.rows else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def timing(lib):
    secs, pts, launches = ctypes.c_double(), ctypes.c_double(), ctypes.c_int64()
    lib.opesci_b200_last_timing(ctypes.byref(secs), ctypes.byref(pts), ctypes.byref(launches))
    return secs.value, pts.value, launches.value


def build_grid(n, nx, steps, warmup, flags):
    import eigenwave3d as drv
    dt = 0.25 / n   # inside the CFL limit 0.495*dx/vp (staggeredgrid.py:99-100)
    g = drv.eigenwave3d((nx / float(n), 1.0, 1.0), (nx, n, n), dt, dt * (steps + warmup), accuracy_order=[2, 4, 4, 4],
                        o_converge=True, verbose=False)
    g.ntsteps.value = steps + warmup
    g.b200_flags = flags
    return g


def cpu_reference_sample(kind_tag="n512"):
    """Time-loop throughput of the reference's generated code on the host cores, by differencing
    two prebuilt step counts of the same grid (oracle/_ref, built from /root/reference)."""
    man_path = os.path.join(ROOT, "oracle", "_ref", "manifest.json")
    if not os.path.exists(man_path):
        return None
    man = json.load(open(man_path))
    pair = sorted((c for name, c in man.items() if name.startswith("ew_bench_so4_f32_" + kind_tag)),
                  key=lambda c: c["steps"])
    if len(pair) != 2 or not all(os.path.exists(os.path.join(ROOT, c["exe"])) for c in pair):
        return None
    cores = os.cpu_count() or 1
    env = dict(os.environ, OMP_NUM_THREADS=str(cores), OMP_PROC_BIND="close")
    secs = []
    for c in pair:
        out = subprocess.check_output([os.path.join(ROOT, c["exe"]), "--time"], env=env).decode()
        secs.append(float(re.search(r"EXECUTE_SECONDS (\S+)", out).group(1)))
    dsteps = pair[1]["steps"] - pair[0]["steps"]
    npts = 1.0
    for v in pair[0]["grid_size"]:
        npts *= v + 1
    loop = max(secs[1] - secs[0], 1e-9)
    return {"value": npts * dsteps / loop / 1e9, "unit": "Gpts/s", "cores": cores, "kind": "reference",
            "sample": "reference generated C++ (g++ -O3 -fopenmp), eigenwave3d so=4 fp32 %d^3, %d steps "
                      "(difference of %d- and %d-step runs: %.2fs - %.2fs)"
                      % (pair[0]["grid_size"][0], dsteps, pair[1]["steps"], pair[0]["steps"], secs[1], secs[0])}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    samples = []
    warm = min(args.warmup, 1)
    reps = max(1, min(args.steps, 3))
    base = None
    for i in range(warm + reps):
        base = cpu_reference_sample("n512" if args.n >= 512 else "n256")
        if base is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref binaries missing (run oracle/refgen/make_ref.py)"}))
            return 0
        if i >= warm:
            samples.append(base["value"])
    value = statistics.median(samples)
    base["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "Gpts/s", "n_gpus": args.gpus,
            "steps": reps, "warmup": warm,
            # time one time step of the named workload takes at the measured rate (the timed sample itself is 512^3)
            "ms_per_step": float(args.n + 1) ** 3 * args.gpus / (value * 1e9) * 1e3,
            "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "eigenwave3d so=4 fp32 %d^3 (reference generated OpenMP C++ on host cores; "
                                   "bounded sample at 512^3)" % args.n},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": "Gpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--n", type=int, default=1024, help="grid cells per axis (BASELINE config: 1024)")
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--arith", default=os.environ.get("OPESCI_B200_ARITH", "fast"), choices=("fast", "reference"))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    torch.cuda.set_device(local_rank)
    from opesci_fd_b200 import abi
    lib = abi.load_library()          # fails loudly if the CUDA library is missing
    if world > 1:
        # x-slab decomposition: the library exchanges halos with NCCL; the 128-byte id travels over torch.distributed
        ident = torch.zeros(abi.COMM_ID_BYTES, dtype=torch.uint8, device="cuda")
        if rank == 0:
            buf = (ctypes.c_ubyte * abi.COMM_ID_BYTES)()
            if lib.opesci_b200_comm_unique_id(buf, abi.COMM_ID_BYTES) != 0:
                raise RuntimeError(lib.opesci_b200_last_error().decode())
            ident = torch.tensor(list(buf), dtype=torch.uint8, device="cuda")
        dist.broadcast(ident, 0)
        buf = (ctypes.c_ubyte * abi.COMM_ID_BYTES)(*ident.cpu().tolist())
        if lib.opesci_b200_comm_init(rank, world, buf, abi.COMM_ID_BYTES) != 0:
            raise RuntimeError(lib.opesci_b200_last_error().decode())
    arith = abi.ARITH_FAST if args.arith == "fast" else abi.ARITH_REFERENCE
    steps, warmup = args.steps, max(args.warmup, 3)
    n = args.n

    # ---- value: device-resident fields, K timed steps after W warm-up steps
    # weak scaling: every GPU owns n planes of an (n*world) x n x n grid (slabs along x, the slowest axis)
    grid = build_grid(n, n * world, steps, warmup, arith | abi.HOST_MIRROR_NONE)
    params, keep = grid.build_params()
    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    grid._library = lib
    orig_build = grid.build_params

    def with_warmup():
        p, k = orig_build()
        p.warmup_steps = warmup
        p.slab_rank, p.slab_nranks = rank, world
        return p, k
    grid.build_params = with_warmup
    grid.run(library=lib)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.stop_flag = True
    sampler.join()
    loop_s, pts, launches = timing(lib)
    if world > 1:
        t = torch.tensor([loop_s], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        loop_s = float(t.item())
    value = pts * steps / loop_s / 1e9          # pts = interior points of the GLOBAL grid
    pts_gpu = pts / world
    # dominant kernel, timed live with events on its launching stream
    kms = (ctypes.c_double * 3)()
    if lib.opesci_b200_time_kernels(ctypes.byref(grid._arg_grid), 5, kms) != 0:
        raise RuntimeError(lib.opesci_b200_last_error().decode())
    l2 = grid.convergence_f64()
    grid.free()
    peak, peak_src = measured_peak()
    fused = kms[1] == 0.0
    step_ms = kms[0] + kms[1] + kms[2]
    if fused:
        dom_name, dom_ms, dom_bytes = "fused stress+velocity", kms[0], BYTES_PER_POINT * pts_gpu
    else:
        # two-pass path: the stress kernel dominates; its own compulsory traffic is 9 reads + 6 writes
        dom_name, dom_ms, dom_bytes = "stress_interior (two-pass path: 15 words/pt)", kms[0], 60.0 * pts_gpu
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    # DRAM bytes of one launch of the dominant kernel from the committed `ncu --set full` capture
    # (profiles/r01b_fused_ncu_summary.txt: dram__bytes_read.sum + dram__bytes_write.sum); only known for
    # the configuration that capture was taken on
    traffic = None
    if fused and world == 1 and n == 1024:
        traffic = 50.320062e9 + 39.194970e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "algorithmic_bytes_per_launch": dom_bytes,
                "kernel": dom_name, "kernel_ms": dom_ms, "peak_source": peak_src,
                "step_algorithmic_GBps": value / world * BYTES_PER_POINT,
                "step_frac_of_peak": value / world * BYTES_PER_POINT / peak,
                "kernel_ms_breakdown": {"stress_or_fused": kms[0], "velocity": kms[1], "ghost_loops": kms[2]},
                "kernel_share_of_step": dom_ms / step_ms if step_ms > 0 else None}

    # ---- e2e: reference-facing ABI call with host result arrays (rank 0 describes its own call)
    e2e = None
    if not args.no_e2e:
        # N = 1: the reference ABI (host result arrays, all 18 level arrays copied back; the page-locked result pool
        # is reserved beforehand, as the contract's "pinned host memory").  N > 1: the slabs stay on the devices
        # (8 x 80 GB would not fit the host) and the result read back is the L2 metric.
        mirror = abi.HOST_MIRROR_FULL if world == 1 else abi.HOST_MIRROR_NONE
        if world == 1:
            nbytes = 2 * 4 * params.dim[0] * params.dim[1] * params.dim[2]
            if lib.opesci_b200_reserve_host(nbytes, 9) != 0:
                raise RuntimeError(lib.opesci_b200_last_error().decode())
        g2 = build_grid(n, n * world, steps, 0, arith | mirror)
        orig2 = g2.build_params

        def with_slab():
            p, k = orig2()
            p.slab_rank, p.slab_nranks = rank, world
            return p, k
        g2.build_params = with_slab
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g2.run(library=lib)
        conv = g2.convergence()
        wall = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([wall], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wall = float(t.item())
        level_bytes = 4.0 * params.dim[0] * params.dim[1] * params.dim[2]
        e2e = {"value": pts * steps / wall / 1e9, "unit": "Gpts/s",
               "h2d_bytes_per_step": float(ctypes.sizeof(abi.OpesciB200Params)) / steps,
               "d2h_bytes_per_step": (18.0 * level_bytes / steps) if world == 1 else 72.0 / steps, "wall_s": wall,
               "what": ("opesci_b200_configure + opesci_execute (device alloc, init, %d steps, D2H of 9 fields x 2 levels "
                        "into pre-reserved pinned host arrays) + opesci_convergence" % steps) if world == 1 else
                       ("opesci_b200_configure + opesci_execute (device alloc, init, %d steps, slabs stay device-resident) "
                        "+ opesci_convergence (all-reduced L2 norms read back)" % steps),
               "l2_U": conv["U_l2"]}
        g2.free()
        lib.opesci_b200_release_host()

    cpu = None if (args.no_cpu or rank != 0) else cpu_reference_sample("n512")
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "Gpts/s", "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": loop_s / steps * 1e3, "higher_is_better": True,
                "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": ("eigenwave3d so=4 fp32 %dx%dx%d grid, " % (n * world, n, n))
                                       + ("%d^3 per GPU as x-slabs with 8 halo planes exchanged by NCCL send/recv (overlapped with "
                                          "the next step), " % n if world > 1 else "")
                                       + "homogeneous medium, six free surfaces (Levander)",
                           "arithmetic": args.arith, "l2_flush": "no explicit flush: working set 18 x %.2f GB per GPU >> 126 MB L2" % (4e-9 * params.dim[1] ** 3),
                           "l2_U_after_run": l2[0]},
                "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": int(launches),
                "roofline": roofline, "cpu_baseline": cpu}
        print(json.dumps(line))
    if world > 1:
        lib.opesci_b200_comm_finalize()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    # stdout carries exactly ONE line (the JSON record); everything else the run prints goes to stderr
    # (redirected at the file-descriptor level: NCCL prints its version banner with printf)
    sys.stdout.flush()
    _real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    _orig_print = print

    def print(*a, **k):   # noqa: A001 -- the JSON line is the only print() in this file
        k.setdefault("file", _real_stdout)
        _orig_print(*a, **k)
        _real_stdout.flush()
    sys.exit(main())
