#!/usr/bin/env python
"""Generate the committed golden fixtures from the reference's own generated code.

Runs in the development container only (needs oracle/_ref built by oracle/refgen/make_ref.py,
which needs /root/reference).  For every `small` configuration it runs the reference binary
in a fresh process, dumps all fields (all time levels) and stores them losslessly:

  tests/golden/<name>.npz      fields [nfields][nlevels][dim1][dim2][dim3] (raw bits) + printed L2 norms
  tests/golden/norms.json      printed L2 norms of the `default` / `mid` configurations
  tests/golden/hashes.json     sha256 of the raw bits of every field of the `heteromid` configurations
                               (fields too large to commit)
  tests/golden/literals.json   every float literal of the interior kernels as printed in the
                               generated source (pins the host front end's coefficient tables)
"""
import hashlib
import json
import os
import re
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
MAN = json.load(open(os.path.join(ROOT, "oracle", "_ref", "manifest.json")))


def run(cfg, dump=None):
    cmd = [os.path.join(ROOT, cfg["exe"])]
    if dump:
        nelem = cfg["nlevels"] * int(np.prod(cfg["dim"]))
        cmd += ["--dump", dump, str(nelem)]
    env = dict(os.environ, OMP_NUM_THREADS=str(os.cpu_count() or 1))
    out = subprocess.check_output(cmd, env=env, cwd=ROOT).decode()
    return [float(m.group(1)) for m in re.finditer(r"^L2\[\d+\] \S+ (\S+)$", out, re.M)], \
           [m.group(1) for m in re.finditer(r"^L2\[\d+\] (\S+) \S+$", out, re.M)]


def kernel_literals(cfg):
    """{lhs field: [signed literal per term, in emitted order]} of the time-loop interior kernels."""
    src = open(os.path.join(ROOT, cfg["cpp"])).read()
    body = src[src.index("_ti < ntsteps"):]
    out = {}
    for f in cfg["fields"]:
        lvl = "_t2" if cfg["kind"] in ("simplewave3d", "regular_generic") else "_t1"
        m = re.search(r"^\s*%s\[%s\]\[x\]\[y\]\[z\] = (.*);$" % (f, lvl), body, re.M)
        terms = re.findall(r"(?:^|\s)([+-]?)\s*([0-9.]+(?:e[+-]?\d+)?)F\*", m.group(1))
        out[f] = [("-" if s == "-" else "") + lit for s, lit in terms]
    return out


def main():
    # `make_golden.py name-substring ...` regenerates only those entries and merges them into the json files
    only = sys.argv[1:]
    norms, literals, hashes = {}, {}, {}
    if only:
        hashes = json.load(open(os.path.join(HERE, "hashes.json")))
        norms = json.load(open(os.path.join(HERE, "norms.json")))
        literals = json.load(open(os.path.join(HERE, "literals.json")))
    for name, cfg in sorted(MAN.items()):
        tags = set(cfg["tags"])
        if only and not any(o in name for o in only):
            continue
        if "heteromid" in tags or "large" in tags:
            dump = os.path.join(os.environ.get("GOLDEN_TMP", "/tmp"), "golden_%s.bin" % name)
            vals, printed = run(cfg, dump)
            arr = np.memmap(dump, dtype=np.float32, mode="r").reshape(len(cfg["fields"]), cfg["nlevels"], *cfg["dim"])
            keys = ("kind", "so", "grid_size", "dt", "steps", "double", "domain", "dim", "fields") + \
                (("seed",) if "seed" in cfg else ("rho", "vp", "vs"))
            hashes[name] = dict(config={k: cfg[k] for k in keys},
                                sha256=[hashlib.sha256(arr[k]).hexdigest() for k in range(arr.shape[0])],
                                absmax=[float(np.abs(arr[k]).max()) for k in range(arr.shape[0])])
            if "large" in tags:   # these run with converge=True: the printed norms are golden too
                hashes[name].update(l2=vals, l2_printed=printed)
            del arr
            os.remove(dump)
            print("hashes", name, cfg["dim"], printed[:2])
        elif "small" in tags:
            dump = "/tmp/golden_%s.bin" % name
            vals, printed = run(cfg, dump)
            dt = np.float64 if cfg["double"] else np.float32
            arr = np.fromfile(dump, dtype=dt).reshape(len(cfg["fields"]), cfg["nlevels"], *cfg["dim"])
            os.remove(dump)
            np.savez_compressed(os.path.join(HERE, name + ".npz"), fields=arr,
                                l2=np.array(vals, dtype=np.float64), config=json.dumps(cfg, sort_keys=True))
            literals[name] = kernel_literals(cfg)
            print("golden", name, arr.shape, printed[:2])
        elif tags & {"default", "mid"}:
            vals, printed = run(cfg)
            norms[name] = dict(config={k: cfg[k] for k in ("kind", "so", "grid_size", "dt", "steps", "double",
                                                           "domain", "dim", "fields") + (("pde",) if "pde" in cfg else ())},
                               l2=vals, l2_printed=printed)
            if "rho" in cfg:
                norms[name]["config"].update(rho=cfg["rho"], vp=cfg["vp"], vs=cfg["vs"])
            literals[name] = kernel_literals(cfg)
            print("norms", name, printed[:2])
    # generic PDE systems: the assignments of the time loop and of the second initialisation, as emitted
    gk_path = os.path.join(HERE, "generic_kernels.json")
    gk = json.load(open(gk_path)) if os.path.exists(gk_path) else {}
    for name, cfg in sorted(MAN.items()):
        if cfg["kind"] != "regular_generic" or (only and not any(o in name for o in only)):
            continue
        lines = [l.strip() for l in open(os.path.join(ROOT, cfg["cpp"]))]
        step = [l for l in lines if re.match(r"^\w+\[_t2\]\[x\]\[y\]\[z\] = ", l)]
        init2 = []
        for l in lines:
            if re.match(r"^\w+\[_t1\]\[_x\]\[_y\]\[_z\] = ", l) and l not in init2:
                init2.append(l)
        gk[name] = dict(config={k: cfg[k] for k in ("kind", "pde", "so", "grid_size", "dt", "steps", "double", "domain", "dim", "fields")},
                        step=step, init2=init2)
    json.dump(gk, open(gk_path, "w"), indent=1, sort_keys=True)
    json.dump(hashes, open(os.path.join(HERE, "hashes.json"), "w"), indent=1, sort_keys=True)
    json.dump(norms, open(os.path.join(HERE, "norms.json"), "w"), indent=1, sort_keys=True)
    json.dump(literals, open(os.path.join(HERE, "literals.json"), "w"), indent=1, sort_keys=True)


if __name__ == "__main__":
    sys.exit(main())
