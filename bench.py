#!/usr/bin/env python
"""bench.py -- grid-point updates per second of the opesci-fd time loop on B200.

Metric (BASELINE.json / SURVEY.md 8d): Gpts/s = interior points * steps / t_loop / 1e9; one grid-point update = all
fields advanced one time step at one interior point.  The default workload is the configuration the metric is quoted
on (BASELINE.json configs[2]: eigenwave3d so=4 fp32 1024^3); the other BASELINE configurations can be selected:

  python bench.py [--gpus N] [--steps K] [--warmup W] [--config NAME] [--n CELLS] [--arith fast|reference]
  python bench.py --impl reference ...      the reference's own generated OpenMP C++ on the host cores

  --config default      eigenwave3d so=4 fp32, 1024^3 per GPU                      72 B/pt   (BASELINE config 3)
           acoustic512  simplewaveequation regular grid so=4 fp32, 512^3           12 B/pt   (config 2)
           so8 | so12   eigenwave3d so=8 / so=12 fp32, 768^3                       72 B/pt   (config 4)
           so4f64 | so8f64 | so12f64   the same in fp64, 768^3                    144 B/pt   (config 4)
           hetero       eigenwave3d `read` mode, synthetic rho/vp/vs, so=4 fp32,
                        1024^3 per GPU (2048x1024x1024 at --gpus 2)               104 B/pt   (config 5)

One "step" is one time step over the whole grid.  `value` is measured with the fields resident in HBM (CUDA events
around exactly K steps after W warm-up steps, inside the library, on the launching stream; the working set is far
larger than L2).  `e2e` is the same metric through the reference-facing C ABI call `opesci_execute` with HOST result
arrays: allocation, (heterogeneous: the H2D copy of rho, vp, vs), initialisation, K steps and the device->host copy of
every level array are inside its timed region.  `roofline` is computed for the dominant kernel, re-timed live with CUDA
events, from SURVEY.md 8d's algorithmic bytes per point update.  `cpu_baseline` times the reference's generated code
(oracle/_ref, g++ -O3 -fopenmp, and the reference's own flags beside it) on this box's host cores on a bounded sample.
"""
import argparse
import ctypes
import hashlib
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

# bytes_pt: SURVEY.md 8d algorithmic bytes per grid-point update; two_pass_words: words moved per point by the dominant
# kernel of the two-pass path (stress: 3 + 6 reads, 6 writes)
CONFIGS = {
    "default": dict(kind="eigenwave3d", so=4, double=False, n=1024, bytes_pt=72.0, dt_n=0.25, metric=METRIC,
                    what="eigenwave3d so=4 fp32, homogeneous medium, six free surfaces (Levander)"),
    "acoustic512": dict(kind="simplewave3d", so=4, double=False, n=512, bytes_pt=12.0, dt_n=0.128,
                        metric="Gpts/s acoustic simplewaveequation so=4 fp32",
                        what="simplewaveequation regular-grid acoustic propagator so=4 fp32 (3 time levels, no BC)"),
    "so8": dict(kind="eigenwave3d", so=8, double=False, n=768, bytes_pt=72.0, dt_n=0.192,
                metric="Gpts/s elastic eigenwave3d so=8 fp32", what="eigenwave3d so=8 fp32 (Robertsson free surfaces)"),
    "so12": dict(kind="eigenwave3d", so=12, double=False, n=768, bytes_pt=72.0, dt_n=0.192,
                 metric="Gpts/s elastic eigenwave3d so=12 fp32", what="eigenwave3d so=12 fp32 (Robertsson free surfaces)"),
    "so4f64": dict(kind="eigenwave3d", so=4, double=True, n=768, bytes_pt=144.0, dt_n=0.192,
                   metric="Gpts/s elastic eigenwave3d so=4 fp64", what="eigenwave3d so=4 fp64 (Levander)"),
    "so8f64": dict(kind="eigenwave3d", so=8, double=True, n=768, bytes_pt=144.0, dt_n=0.192,
                   metric="Gpts/s elastic eigenwave3d so=8 fp64", what="eigenwave3d so=8 fp64"),
    "so12f64": dict(kind="eigenwave3d", so=12, double=True, n=768, bytes_pt=144.0, dt_n=0.192,
                    metric="Gpts/s elastic eigenwave3d so=12 fp64", what="eigenwave3d so=12 fp64"),
    "hetero": dict(kind="eigenwave3d_read", so=4, double=False, n=1024, bytes_pt=104.0, dt_n=0.4 / 1.5,
                   metric="Gpts/s elastic eigenwave3d-read heterogeneous so=4 fp32",
                   what="eigenwave3d `read` mode so=4 fp32, synthetic random rho/vp/vs per cell (Philox), Levander"),
}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def kernel_source_hash():
    """sha256 over the CUDA sources: keys profiles/traffic.json (DRAM bytes from `ncu --set full` captures), so a
    number captured on one build is never reported for another."""
    h = hashlib.sha256()
    csrc = os.path.join(ROOT, "opesci_fd_b200", "csrc")
    for f in sorted(os.listdir(csrc)):
        if f.endswith((".cu", ".cuh")):
            h.update(open(os.path.join(csrc, f), "rb").read())
    return h.hexdigest()[:16]


def recorded_traffic(config, n, world):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, from the ncu capture
    recorded in profiles/traffic.json for THIS build of the kernels, else None."""
    path = os.path.join(ROOT, "profiles", "traffic.json")
    if world != 1 or not os.path.exists(path):
        return None, None
    key = "%s_n%d" % (config, n)
    ent = json.load(open(path)).get(key)
    if not ent:
        return None, None
    if ent.get("source_hash") != kernel_source_hash():
        return None, "profiles/traffic.json holds %s for another build of the kernels (hash %s)" % (key, ent.get("source_hash"))
    return float(ent["dram_bytes_read"]) + float(ent["dram_bytes_write"]), ent.get("capture")


def mem_available_gb():
    try:
        for line in open("/proc/meminfo"):
            if line.startswith("MemAvailable:"):
                return float(line.split()[1]) / 1e6
    except OSError:
        pass
    return 0.0


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


def build_grid(C, n, world, steps, warmup, flags, lib=None, rank=0, media=True):
    """Front-end grid of configuration C on an (n*world) x n x n grid (weak scaling: x-slabs of n planes per GPU)."""
    from opesci_fd_b200 import abi
    nx = n * world
    dt = C["dt_n"] / n
    order = [2, C["so"], C["so"], C["so"]]
    total = steps + warmup
    domain = (nx / float(n), 1.0, 1.0)
    if C["kind"] == "simplewave3d":
        import simplewaveequation as drv
        g = drv.simplewave3d(domain, (nx, n, n), dt, dt * total, accuracy_order=order, o_converge=True,
                             double=C["double"], verbose=False)
    elif C["kind"] == "eigenwave3d_read":
        import eigenwave3d as drv
        from opesci_fd_b200.util import synthetic_media_planes
        g = drv.eigenwave3d(domain, (nx, n, n), dt, dt * total, accuracy_order=order, o_converge=False, read=True,
                            rho_file="-", vp_file="-", vs_file="-", verbose=False)
        if media:
            dims = [d.value for d in g.dim]
            l0, l1 = ctypes.c_int(0), ctypes.c_int(dims[0])
            if world > 1 and lib.opesci_b200_slab_range(rank, world, dims[0], C["so"], ctypes.byref(l0), ctypes.byref(l1)) != 0:
                raise RuntimeError(lib.opesci_b200_last_error().decode())
            g.set_media_arrays(*synthetic_media_planes(dims, l0.value, l1.value - l0.value), plane0=l0.value)
    else:
        import eigenwave3d as drv
        g = drv.eigenwave3d(domain, (nx, n, n), dt, dt * total, accuracy_order=order, o_converge=True,
                            double=C["double"], verbose=False)
    g.ntsteps.value = total
    g.b200_flags = flags
    orig = g.build_params

    def with_slab():
        p, k = orig()
        p.warmup_steps = warmup
        p.slab_rank, p.slab_nranks = rank, world
        return p, k
    g.build_params = with_slab
    return g


# ---------------------------------------------------------------------------------------------- reference (CPU) arm
def _manifest():
    path = os.path.join(ROOT, "oracle", "_ref", "manifest.json")
    return json.load(open(path)) if os.path.exists(path) else {}


def cpu_reference_sample(tag="n512", exe_key="exe", prefix="ew_bench_so4_f32_"):
    """Time-loop throughput of the reference's generated code on the host cores, by differencing two prebuilt step
    counts of the same grid (oracle/_ref, built from /root/reference by oracle/refgen/make_ref.py)."""
    man = _manifest()
    pair = sorted((c for name, c in man.items() if name.startswith(prefix + tag + "_s") and exe_key in c),
                  key=lambda c: c["steps"])
    if len(pair) != 2 or not all(os.path.exists(os.path.join(ROOT, c[exe_key])) for c in pair):
        return None
    cores = os.cpu_count() or 1
    env = dict(os.environ, OMP_NUM_THREADS=str(cores), OMP_PROC_BIND="close")
    secs = []
    for c in pair:
        if c["kind"] == "eigenwave3d_read":
            # input files of the `read` mode reference: written here when they did not travel with the snapshot
            paths = [os.path.join(ROOT, m) for m in c["media"]]
            if not all(os.path.exists(q) for q in paths):
                from opesci_fd_b200.util import synthetic_media
                os.makedirs(os.path.dirname(paths[0]), exist_ok=True)
                for q, arr in zip(paths, synthetic_media(c["dim"], c["seed"])):
                    arr.tofile(q)
        out = subprocess.check_output([os.path.join(ROOT, c[exe_key]), "--time"], env=env, cwd=ROOT).decode()
        secs.append(float(re.search(r"EXECUTE_SECONDS (\S+)", out).group(1)))
    dsteps = pair[1]["steps"] - pair[0]["steps"]
    npts = 1.0
    for v in pair[0]["grid_size"]:
        npts *= v + 1
    loop = max(secs[1] - secs[0], 1e-9)
    flags = "g++ -O3 -fopenmp" if exe_key == "exe" else "g++ -g -O3 -fno-tree-vectorize -fopenmp (opesci/compilation.py:58)"
    return {"value": npts * dsteps / loop / 1e9, "unit": "Gpts/s", "cores": cores, "kind": "reference",
            "grid": pair[0]["grid_size"],
            "sample": "reference generated C++ (%s), %s so=%d %s %s grid, %d steps (difference of %d- and %d-step runs: "
                      "%.2fs - %.2fs)" % (flags, pair[0]["kind"], pair[0]["so"], "fp64" if pair[0]["double"] else "fp32",
                                          "x".join(str(v) for v in pair[0]["grid_size"]), dsteps, pair[1]["steps"],
                                          pair[0]["steps"], secs[1], secs[0])}


def reference_pair_for(config, n):
    """(prefix, tag, same_grid, why) of the prebuilt reference pair closest to configuration `config` at n^3."""
    C = CONFIGS[config]
    man = _manifest()
    prefix = {"default": "ew_bench_so4_f32_", "acoustic512": "sw_bench_so4_f32_", "so8": "ew_bench_so8_f32_",
              "so12": "ew_bench_so12_f32_", "so4f64": "ew_bench_so4_f64_", "so8f64": "ew_bench_so8_f64_",
              "so12f64": "ew_bench_so12_f64_", "hetero": "ewh_bench_so4_f32_"}[config]
    sizes = sorted({int(re.search(r"_n(\d+)_s", k).group(1)) for k in man if k.startswith(prefix) and re.search(r"_n(\d+)_s", k)})
    if not sizes:
        return None
    # host memory the reference needs: every level array, dense, plus the media arrays in read mode
    nfields = 1 if C["kind"] == "simplewave3d" else 9
    levels = 3 if C["kind"] == "simplewave3d" else 2
    esz = 8 if C["double"] else 4

    def need_gb(s):
        d = s + 1 + C["so"]
        return (nfields * levels + (13 if C["kind"] == "eigenwave3d_read" else 0)) * esz * d ** 3 / 1e9
    avail = mem_available_gb()
    fit = [s for s in sizes if s <= n and need_gb(s) * 1.15 < avail]
    if not fit:
        return None
    s = max(fit)
    why = None
    if s != n:
        bigger = [t for t in sizes if s < t <= n]
        if bigger:
            why = "%d^3 needs %.0f GB of host memory, %.0f GB available" % (min(bigger), need_gb(min(bigger)), avail)
        else:
            why = "largest prebuilt reference binary for this configuration is %d^3" % s
    return prefix, "n%d" % s, s == n, why


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    C = CONFIGS[args.config]
    n = args.n or C["n"]
    sel = reference_pair_for(args.config, n)
    if sel is None:
        # the reference's own default-size pair always exists when oracle/_ref was built
        sel = ("ew_bench_so4_f32_", "n512", False, "no prebuilt reference pair for configuration %s" % args.config) \
            if args.config == "default" else None
    if sel is None:
        print(json.dumps({"impl": "reference", "unavailable": "no prebuilt oracle/_ref pair for --config %s "
                          "(oracle/refgen/make_ref.py --tags bench)" % args.config}))
        return 0
    prefix, tag, same, why = sel
    big = int(tag[1:]) >= 1000
    warm = 0 if big else min(args.warmup, 1)
    reps = 1 if big else max(1, min(args.steps, 3))
    samples, base = [], None
    for i in range(warm + reps):
        base = cpu_reference_sample(tag, "exe", prefix)
        if base is None:
            print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref binaries missing (run oracle/refgen/make_ref.py)"}))
            return 0
        if i >= warm:
            samples.append(base["value"])
    value = statistics.median(samples)
    base["value"] = value
    pts_step = float(n + 1) ** 3 * args.gpus
    line = {"impl": "reference", "metric": C["metric"], "value": value, "unit": "Gpts/s", "n_gpus": args.gpus,
            "steps": reps, "warmup": warm,
            # the time one step of the named workload takes at the measured rate
            "ms_per_step": pts_step / (value * 1e9) * 1e3, "ms_per_step_is": "workload points / measured rate",
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64" if C["double"] else "f32", "data": "synthetic",
            "config": {"workload": "%s %d^3 -- the reference's own generated OpenMP C++ on the host cores" % (C["what"], n),
                       "sample_grid": base["grid"], "same_config": bool(same), "why_not_same_grid": why},
            "cpu_baseline": base,
            "e2e": {"value": value, "unit": "Gpts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))
    return 0


# ---------------------------------------------------------------------------------------------------- slab parity
def slab_parity_check(lib, rank, world, dist, torch):
    """N > 1: before the timed run, a small N-rank run is compared with the single-domain run of the same grid --
    every owned plane of every field, bit for bit (reference arithmetic).  Returns the verdict string."""
    import numpy as np
    from opesci_fd_b200 import abi
    from common import make_grid
    planes = 40 * world
    cfg = dict(kind="eigenwave3d", so=4, grid_size=[planes, 70, 130], dt=0.002, steps=9, double=False,
               domain=[planes / 64.0, 1.0, 1.0])
    g = make_grid(cfg, flags=abi.ARITH_REFERENCE | abi.HOST_MIRROR_FULL)
    orig = g.build_params

    def with_slab():
        p, k = orig()
        p.slab_rank, p.slab_nranks = rank, world
        return p, k
    g.build_params = with_slab
    g.run(library=lib)
    p = g._params
    dims = [p.dim[0], p.dim[1], p.dim[2]]
    l0, l1 = ctypes.c_int(), ctypes.c_int()
    lib.opesci_b200_slab_range(rank, world, dims[0], 4, ctypes.byref(l0), ctypes.byref(l1))
    nloc = l1.value - l0.value
    mine = []
    for k in range(9):
        buf = ctypes.cast(g._arg_grid.field[k], ctypes.POINTER(ctypes.c_float * (2 * nloc * dims[1] * dims[2]))).contents
        mine.append(np.frombuffer(buf, dtype=np.float32).reshape(2, nloc, dims[1], dims[2]).copy())
    g.free()
    # single-domain run of the same grid on this GPU
    s = make_grid(cfg, flags=abi.ARITH_REFERENCE | abi.HOST_MIRROR_FULL)
    s.run(library=lib)
    bad = 0
    # owned planes: the slab minus its halo planes (physical ghost planes of the end ranks included)
    halo = abi.SLAB_HALO if hasattr(abi, "SLAB_HALO") else 8
    own_lo = 0 if rank == 0 else l0.value + halo
    own_hi = dims[0] if rank == world - 1 else l1.value - halo
    for k in range(9):
        full = s.field_array(k)
        a = mine[k][:, own_lo - l0.value:own_hi - l0.value]
        b = full[:, own_lo:own_hi]
        bad += int((a.view(np.int32) != b.view(np.int32)).sum())
    s.free()
    t = torch.tensor([bad], device="cuda", dtype=torch.int64)
    dist.all_reduce(t)
    return "bit_identical" if int(t.item()) == 0 else "MISMATCH (%d cells)" % int(t.item())


# --------------------------------------------------------------------------------------------------------- main
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--config", default="default", choices=sorted(CONFIGS))
    ap.add_argument("--n", type=int, default=0, help="grid cells per axis and GPU (default: the BASELINE size of --config)")
    ap.add_argument("--impl", default="ours", choices=("ours", "reference"))
    ap.add_argument("--arith", default=os.environ.get("OPESCI_B200_ARITH", "fast"), choices=("fast", "reference"))
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ref-arith", action="store_true", help="skip the second (bit-exact arithmetic) device-resident run")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    C = CONFIGS[args.config]
    n = args.n or C["n"]

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

    slab_parity = None
    if world > 1:
        slab_parity = slab_parity_check(lib, rank, world, dist, torch)

    # ---- value: device-resident fields, K timed steps after W warm-up steps
    # weak scaling: every GPU owns n planes of an (n*world) x n x n grid (slabs along x, the slowest axis)
    t_media0 = time.perf_counter()
    grid = build_grid(C, n, world, steps, warmup, arith | abi.HOST_MIRROR_NONE, lib, rank)
    media_gen_s = time.perf_counter() - t_media0
    params, keep = grid.build_params()
    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    grid.run(library=lib)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    sampler.stop_flag = True
    sampler.join()
    loop_s, pts, launches = timing(lib)
    halo_transport = abi.HALO_TRANSPORTS.get(lib.opesci_b200_halo_transport(), "?") if world > 1 else None
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
    l2 = grid.convergence_f64() if C["kind"] != "eigenwave3d_read" else None
    parts = (ctypes.c_double * 4)()
    if hasattr(lib, "opesci_b200_time_fused_parts") and lib.opesci_b200_time_fused_parts(ctypes.byref(grid._arg_grid), 5, parts) != 0:
        raise RuntimeError(lib.opesci_b200_last_error().decode())
    media_keep = (getattr(grid, "media_arrays", None), getattr(grid, "media_plane0", 0))
    grid.free()
    peak, peak_src = measured_peak()
    bytes_pt = C["bytes_pt"]
    esz = 8 if C["double"] else 4
    step_ms = kms[0] + kms[1] + kms[2]
    if C["kind"] == "simplewave3d":
        dom_name, dom_ms, dom_bytes = "acoustic_march (whole step)", kms[0], bytes_pt * pts_gpu
    elif kms[1] == 0.0 and parts[0] > 0.0:
        # z-fold: the fused step is two concurrent launches.  The dominant one is the interior launch (tile columns
        # 1 .. nzt-2), timed on its own like ncu times it; its algorithmic bytes are those of the z columns it stores.
        dom_name = "fused_step interior launch (%d of %d z columns; the z-edge launch runs beside it)" % (parts[2], parts[3])
        dom_ms, dom_bytes = parts[0], bytes_pt * pts_gpu * parts[2] / parts[3]
    elif kms[1] == 0.0:
        dom_name, dom_ms, dom_bytes = "fused stress+velocity", kms[0], bytes_pt * pts_gpu
    else:
        # two-pass path: the stress kernel dominates; its own compulsory traffic is 3 + 6 reads and 6 writes
        # (+ 5 media words in read mode)
        words = 15 + (5 if C["kind"] == "eigenwave3d_read" else 0)
        dom_name, dom_ms, dom_bytes = "stress_tiled (two-pass path: %d words/pt)" % words, kms[0], words * esz * pts_gpu
    achieved = dom_bytes / (dom_ms * 1e-3) / 1e9
    traffic, traffic_src = recorded_traffic(args.config, n, world)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "algorithmic_bytes_per_launch": dom_bytes,
                "kernel": dom_name, "kernel_ms": dom_ms, "peak_source": peak_src,
                "step_algorithmic_GBps": value / world * bytes_pt,
                "step_frac_of_peak": value / world * bytes_pt / peak,
                "bytes_per_point": bytes_pt,
                "kernel_ms_breakdown": {"stress_or_fused": kms[0], "velocity": kms[1], "ghost_loops": kms[2],
                                        "fused_interior_launch_alone": parts[0] or None, "fused_zedge_launch_alone": parts[1] or None},
                "fused_both_launches": {"ms": kms[0], "achieved": bytes_pt * pts_gpu / (kms[0] * 1e-3) / 1e9,
                                        "frac": bytes_pt * pts_gpu / (kms[0] * 1e-3) / 1e9 / peak} if parts[0] > 0.0 else None,
                "kernel_share_of_step": dom_ms / step_ms if step_ms > 0 else None,
                "kernel_source_hash": kernel_source_hash()}

    # ---- the bit-exact arithmetic mode beside it (the tests pin THIS mode bit for bit on the reference)
    value_ref = None
    if not args.no_ref_arith and args.arith == "fast":
        ksteps = min(steps, 20)
        gr = build_grid(C, n, world, ksteps, 3, abi.ARITH_REFERENCE | abi.HOST_MIRROR_NONE, lib, rank, media=False)
        if media_keep[0] is not None:
            gr.media_arrays, gr.media_plane0 = media_keep
        gr.run(library=lib)
        ls, p2, _ = timing(lib)
        if world > 1:
            t = torch.tensor([ls], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ls = float(t.item())
        value_ref = p2 * ksteps / ls / 1e9
        gr.free()

    # ---- e2e: reference-facing ABI call with host result arrays
    e2e = None
    if not args.no_e2e:
        nlev = params.nlevels
        level_bytes = float(esz) * params.dim[1] * params.dim[2] * (params.dim[0] / world + (0 if world == 1 else 16))
        result_gb = params.nfields * nlev * level_bytes / 1e9
        # every rank copies all level arrays of its slab back into pinned host memory when the host has room for them
        host_ok = mem_available_gb() > 1.4 * result_gb * min(world, 8)     # (never drive the box near its memory limit)
        if world > 1:
            t = torch.tensor([1 if host_ok else 0], device="cuda", dtype=torch.int64)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            host_ok = bool(int(t.item()))
        mirror = abi.HOST_MIRROR_FULL if host_ok else abi.HOST_MIRROR_NONE
        g2 = build_grid(C, n, world, steps, 0, arith | mirror, lib, rank, media=False)
        if media_keep[0] is not None:
            g2.media_arrays, g2.media_plane0 = media_keep
        p2, _k2 = g2.build_params()
        if host_ok:
            nbytes = nlev * esz * p2.dim[1] * p2.dim[2]
            l0, l1 = ctypes.c_int(0), ctypes.c_int(p2.dim[0])
            if world > 1:
                lib.opesci_b200_slab_range(rank, world, p2.dim[0], C["so"], ctypes.byref(l0), ctypes.byref(l1))
            nbytes *= (l1.value - l0.value)
            if lib.opesci_b200_reserve_host(nbytes, params.nfields) != 0:
                raise RuntimeError(lib.opesci_b200_last_error().decode())
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        g2.run(library=lib)
        conv = g2.convergence() if C["kind"] != "eigenwave3d_read" else {}
        wall = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([wall], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wall = float(t.item())
        h2d = float(ctypes.sizeof(abi.OpesciB200Params))
        if C["kind"] == "eigenwave3d_read":
            h2d += 3.0 * level_bytes * (4.0 / esz)
        d2h = params.nfields * nlev * level_bytes if host_ok else 8.0 * params.nfields
        e2e = {"value": pts * steps / wall / 1e9, "unit": "Gpts/s",
               "h2d_bytes_per_step": h2d / steps, "d2h_bytes_per_step": d2h / steps, "wall_s": wall,
               "device_resident": not host_ok,
               "what": ("opesci_b200_configure + opesci_execute (device alloc, %sinit, %d steps, D2H of %d fields x %d levels%s "
                        "into pre-reserved pinned host arrays) + opesci_convergence"
                        % ("H2D of rho/vp/vs, media derivation, " if C["kind"] == "eigenwave3d_read" else "", steps,
                           params.nfields, nlev, " of every rank's slab" if world > 1 else "")) if host_ok else
                       ("opesci_b200_configure + opesci_execute (device alloc, init, %d steps; the result arrays (%.0f GB per rank) "
                        "do not fit the host's available memory, so they stay device-resident) + opesci_convergence "
                        "(L2 norms read back)" % (steps, result_gb)),
               "l2_first_field": (list(conv.values())[0] if conv else None)}
        g2.free()
        lib.opesci_b200_release_host()

    cpu = None
    if not args.no_cpu and rank == 0:
        if args.config == "default":
            cpu = cpu_reference_sample("n512")
            if cpu is not None:
                rf = cpu_reference_sample("n256", "exe_refflags")
                if rf is not None:
                    cpu["reference_flags"] = {k: rf[k] for k in ("value", "unit", "cores", "sample")}
        else:
            sel = reference_pair_for(args.config, n)
            if sel is not None:
                cpu = cpu_reference_sample(sel[1], "exe", sel[0])
    if rank == 0:
        dims_txt = "%dx%dx%d grid" % (n * world, n, n)
        line = {"metric": C["metric"], "value": value, "unit": "Gpts/s", "n_gpus": world, "steps": steps, "warmup": warmup,
                "ms_per_step": loop_s / steps * 1e3, "higher_is_better": True,
                "scaling": "weak",
                "vs_baseline": None, "dtype": "f64" if C["double"] else "f32", "data": "synthetic",
                "config": {"workload": "%s, %s" % (C["what"], dims_txt)
                                       + (", %d^3 per GPU as x-slabs with 8 halo planes refreshed after the stress / the velocity ghost loops "
                                          "(transport: see halo_transport)" % n if world > 1 else ""),
                           "name": args.config,
                           "arithmetic": args.arith,
                           "l2_flush": "no explicit flush: working set %.1f GB per GPU >> 126 MB L2"
                                       % (params.nfields * params.nlevels * esz * 1e-9 * params.dim[1] * params.dim[2] * (params.dim[0] / world)),
                           "l2_first_field_after_run": (l2[0] if l2 else None)},
                "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": int(launches),
                "value_reference_arith": value_ref,
                "roofline": roofline, "cpu_baseline": cpu}
        if slab_parity is not None:
            line["slab_parity"] = slab_parity
        if halo_transport is not None:
            line["config"]["halo_transport"] = halo_transport
        if C["kind"] == "eigenwave3d_read":
            line["config"]["media_generation_s"] = media_gen_s
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
