#!/usr/bin/env python
"""Static SASS instruction counts per kernel of the shipped library (no GPU needed).

    python tools/sass_counts.py [kernel-name regex] > profiles/<round>_sass_evidence.txt

Runs `cuobjdump -sass` on opesci_fd_b200/csrc/libopesci_b200.so and counts, per kernel, the mnemonics that prove what the
kernel uses: UTMALDG / UTMASTG (TMA tensor loads / stores), SYNCS (mbarrier operations), UCGABAR_ARV / _WAIT (cluster barrier),
STAS (st.async into a peer CTA's shared memory), SHFL, LDG / STG / LDS / STS, LDL / STL (register spills) and FFMA / DFMA.
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "opesci_fd_b200", "csrc", "libopesci_b200.so")
COLS = ["UTMALDG", "UTMASTG", "SYNCS", "UCGABAR", "STAS", "BAR", "SHFL", "LDG", "STG", "LDS", "STS", "LDL", "STL", "FFMA", "DFMA"]


def main():
    pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else "fused_step|stress_tiled|velocity_tiled|vel_zface|acoustic_march")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = {}
    counts = collections.defaultdict(collections.Counter)
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,8}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1).split(".")[0]
            counts[cur]["UCGABAR" if op.startswith("UCGABAR") else op] += 1
    mangled = list(counts)
    demangled = subprocess.run(["c++filt"], input="\n".join(mangled), capture_output=True, text=True, check=True).stdout.splitlines()
    for a, b in zip(mangled, demangled):
        b = re.sub(r"^void ", "", b)
        names[a] = re.sub(r"\(.*$", "", b).replace("(int)", "").replace("(bool)", "")
    print("# cuobjdump -sass opesci_fd_b200/csrc/libopesci_b200.so (sm_100a), static instruction counts per kernel (tools/sass_counts.py)")
    print("# UTMALDG / UTMASTG = TMA tensor load / store, SYNCS = mbarrier operations, UCGABAR = cluster barrier, STAS = st.async into the")
    print("# partner CTA's shared memory (distributed shared memory), LDL / STL = local-memory (spill) traffic.")
    print("# fused_step<SO, ARITH (0 reference / 1 fast), HET, ZF (z-edge tiles), PAIR (2-CTA clusters)>")
    print("%-58s" % "kernel" + "".join("%8s" % c for c in COLS))
    for k in sorted(mangled, key=lambda x: names[x]):
        if not pat.search(names[k]):
            continue
        print("%-58s" % names[k][:57] + "".join("%8d" % counts[k][c] for c in COLS))


if __name__ == "__main__":
    main()
