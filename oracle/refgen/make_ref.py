#!/usr/bin/env python
"""Build `oracle/_ref/`: the reference's OWN generated OpenMP C++ for a list of configurations.

TEST INFRASTRUCTURE ONLY.  Runs in the development container (needs /root/reference, sympy);
its outputs (generated .cpp, binaries, manifest.json) are git-ignored but travel to the GPU
box with the gpurun snapshot, where `bench.py --impl reference` and the parity tests use the
prebuilt binaries only.

Pipeline per configuration:
  shim_reference.py  -> scratch py3 copy of /root/reference/opesci (mechanical edits only)
  tests/eigenwave3d.py:eigenwave3d() / tests/simplewaveequation.py:simplewave3d()
                     -> grid object  -> grid.generate()            (reference code path)
  g++ -O3 -fopenmp   -> oracle/_ref/bin/<name>   (wrapper ref_main.cpp + zero_alloc.c)

Two step counts (`ntsteps`) can be generated for one configuration so that the time loop
can be timed by differencing (init + L2 are outside the metric).
"""
import argparse
import concurrent.futures as cf
import json
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
OUT = os.path.join(ROOT, "oracle", "_ref")
SHIM_DIR = os.environ.get("OPESCI_SHIM_DIR", "/tmp/opesci_py3")

# name -> config.  `steps` overrides the reference's int(tmax/dt) so float division
# never decides the step count.  rho/vp/vs go through grid.set_media_params (reference API).
CONFIGS = {}


def _ew(name, so, grid_size, dt, steps, double=False, domain=(1.0, 1.0, 1.0),
        rho=1.0, vp=1.0, vs=0.5, converge=True, tags=()):
    CONFIGS[name] = dict(kind="eigenwave3d", so=so, grid_size=list(grid_size), dt=dt, steps=steps,
                         double=double, domain=list(domain), rho=rho, vp=vp, vs=vs,
                         converge=converge, tags=list(tags))


def _sw(name, so, grid_size, dt, steps, double=False, domain=(1.0, 1.0, 1.0), tags=()):
    CONFIGS[name] = dict(kind="simplewave3d", so=so, grid_size=list(grid_size), dt=dt, steps=steps,
                         double=double, domain=list(domain), converge=True, tags=list(tags))


# --- the reference's own default test (tests/eigenwave3d.py:149-167), every order: golden L2 table
for _so in (2, 4, 6, 8, 10, 12):
    _ew("ew_default_so%d_f32" % _so, _so, (100, 100, 100), 0.002, 500, tags=["default"])
_ew("ew_default_so4_f64", 4, (100, 100, 100), 0.002, 500, double=True, tags=["default"])
_sw("sw_default_so4_f32", 4, (100, 100, 100), 0.002, 500, tags=["default"])
# --- small anisotropic cases with non-trivial media: full-field parity fixtures (bit-exact pinning)
for _so in (2, 4, 6, 8, 10, 12):
    _ew("ew_small_so%d_f32" % _so, _so, (14, 12, 10), 0.004, 7, domain=(1.0, 0.9, 0.8),
        rho=1.3, vp=1.7, vs=0.9, tags=["small"])
for _so in (4, 8, 12):
    _ew("ew_small_so%d_f64" % _so, _so, (14, 12, 10), 0.004, 7, double=True, domain=(1.0, 0.9, 0.8),
        rho=1.3, vp=1.7, vs=0.9, tags=["small"])
_ew("ew_small_so4_f32_even", 4, (14, 12, 10), 0.004, 8, domain=(1.0, 0.9, 0.8),
    rho=1.3, vp=1.7, vs=0.9, tags=["small"])
for _so in (2, 4, 8):
    _sw("sw_small_so%d_f32" % _so, _so, (14, 12, 10), 0.002, 7, domain=(1.0, 0.9, 0.8), tags=["small"])
_sw("sw_small_so4_f64", 4, (14, 12, 10), 0.002, 7, double=True, domain=(1.0, 0.9, 0.8), tags=["small"])
# --- medium: 64^3, enough steps for rounding differences to show
_ew("ew_mid_so4_f32", 4, (64, 64, 64), 0.003, 60, tags=["mid"])
_ew("ew_mid_so8_f32", 8, (64, 64, 64), 0.003, 60, tags=["mid"])
_ew("ew_mid_so4_f64", 4, (64, 64, 64), 0.003, 60, double=True, tags=["mid"])
# --- CPU-baseline timing pairs (bench.py --impl reference): same grid, two step counts
_ew("ew_bench_so4_f32_n256_s4", 4, (256, 256, 256), 0.001, 4, converge=False, tags=["bench"])
_ew("ew_bench_so4_f32_n256_s84", 4, (256, 256, 256), 0.001, 84, converge=False, tags=["bench"])
_ew("ew_bench_so4_f32_n512_s2", 4, (512, 512, 512), 0.0005, 2, converge=False, tags=["bench"])
_ew("ew_bench_so4_f32_n512_s34", 4, (512, 512, 512), 0.0005, 34, converge=False, tags=["bench"])


def generate(name):
    """Runs in a fresh process: build the grid through the reference API and emit C++."""
    cfg = CONFIGS[name]
    sys.path.insert(0, SHIM_DIR)
    sys.path.insert(0, os.path.join(SHIM_DIR, "drivers"))
    t0 = time.time()
    so = cfg["so"]
    order = [2, so, so, so]
    dt, steps = cfg["dt"], cfg["steps"]
    gen_dir = os.path.join(OUT, "gen")
    os.makedirs(gen_dir, exist_ok=True)
    cpp = os.path.join(gen_dir, name + ".cpp")
    devnull = open(os.devnull, "w")
    stdout, sys.stdout = sys.stdout, devnull
    try:
        if cfg["kind"] == "eigenwave3d":
            import eigenwave3d as drv
            grid = drv.eigenwave3d(tuple(cfg["domain"]), tuple(cfg["grid_size"]), dt, dt * steps,
                                   accuracy_order=order, o_converge=cfg["converge"], omp=True,
                                   simd=False, ivdep=True, double=cfg["double"], filename=cpp)
            grid.set_media_params(read=False, rho=cfg["rho"], vp=cfg["vp"], vs=cfg["vs"])
        else:
            import simplewaveequation as drv
            grid = drv.simplewave3d(tuple(cfg["domain"]), tuple(cfg["grid_size"]), dt, dt * steps,
                                    accuracy_order=order, o_converge=cfg["converge"], omp=True,
                                    simd=False, ivdep=True, double=cfg["double"], filename=cpp)
        grid.ntsteps.value = steps
        grid.generate(cpp)
    finally:
        sys.stdout = stdout
    info = dict(cfg)
    info.update(name=name, cpp=os.path.relpath(cpp, ROOT),
                dim=[int(d.value) for d in grid.dim], margin=int(grid.margin.value),
                dx=[float(s.value) for s in grid.spacing], nlevels=len(grid.time),
                fields=[str(f.label) for f in grid.fields], gen_seconds=round(time.time() - t0, 1))
    if cfg["kind"] == "eigenwave3d":
        dv = grid.defined_variable
        info.update({"lambda": float(dv["lambda"].value), "mu": float(dv["mu"].value),
                     "beta": float(dv["beta"].value)})
    return info


def build(info, cxxflags=("-O3", "-fopenmp"), suffix=""):
    bin_dir = os.path.join(OUT, "bin")
    os.makedirs(bin_dir, exist_ok=True)
    exe = os.path.join(bin_dir, info["name"] + suffix)
    real_t = "double" if info["double"] else "float"
    zobj = os.path.join(bin_dir, "zero_alloc.o")
    if not os.path.exists(zobj):
        subprocess.check_call(["gcc", "-O2", "-c", os.path.join(HERE, "zero_alloc.c"), "-o", zobj])
    cmd = ["g++", "-std=c++11", "-w"] + list(cxxflags) + [
        '-DOPESCI_GENERATED="%s"' % os.path.join(ROOT, info["cpp"]), "-DOPESCI_REAL_T=%s" % real_t,
        os.path.join(HERE, "ref_main.cpp"), zobj, "-o", exe]
    subprocess.check_call(cmd)
    return os.path.relpath(exe, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="", help="comma-separated substrings of config names")
    ap.add_argument("--tags", default="", help="comma-separated tags (default, small, mid, bench)")
    ap.add_argument("--jobs", type=int, default=max(1, (os.cpu_count() or 2) - 1))
    ap.add_argument("--refflags", action="store_true",
                    help="also build with the reference's own flags (opesci/compilation.py:58)")
    args = ap.parse_args()
    if not os.path.isdir(os.path.join(SHIM_DIR, "opesci")):
        sys.path.insert(0, HERE)
        import shim_reference
        shim_reference.shim(SHIM_DIR)
    names = list(CONFIGS)
    if args.only:
        pats = args.only.split(",")
        names = [n for n in names if any(p in n for p in pats)]
    if args.tags:
        tags = set(args.tags.split(","))
        names = [n for n in names if tags & set(CONFIGS[n]["tags"])]
    os.makedirs(OUT, exist_ok=True)
    man_path = os.path.join(OUT, "manifest.json")
    manifest = {}
    if os.path.exists(man_path):
        with open(man_path) as fh:
            manifest = json.load(fh)
    with cf.ProcessPoolExecutor(max_workers=args.jobs, max_tasks_per_child=1) as ex:
        futs = {ex.submit(generate, n): n for n in names}
        for fut in cf.as_completed(futs):
            info = fut.result()
            info["exe"] = build(info)
            if args.refflags:
                info["exe_refflags"] = build(info, ("-g", "-O3", "-fno-tree-vectorize", "-fopenmp"),
                                             "_refflags")
            manifest[info["name"]] = info
            print("built %-32s dim=%s gen=%ss" % (info["name"], info["dim"], info["gen_seconds"]), flush=True)
            with open(man_path, "w") as fh:
                json.dump(manifest, fh, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
