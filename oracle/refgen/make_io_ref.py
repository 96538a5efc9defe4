#!/usr/bin/env python
"""Build oracle/_ref/libopesci_io_ref.so: the reference's own libopesci I/O helpers.

TEST INFRASTRUCTURE ONLY.  Compiles /root/reference/src/opesciIO.cpp and opesciHandy.cpp where they lie
(g++ on the two files, no cmake, no VTK: the VTK writers sit behind `#ifdef VTK_FOUND`) together with the
extern "C" doors of io_ref_wrap.cpp.  `-include cmath`: opesciIO.cpp uses pow/fabs without including
<cmath> (it relied on a transitive include of older libstdc++).  Development container only; the GPU box
uses the prebuilt file and the committed fixtures (tests/golden/io_golden.npz).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "_ref")
REF = "/root/reference"


def build():
    if not os.path.isdir(os.path.join(REF, "src")):
        return None
    os.makedirs(OUT, exist_ok=True)
    lib = os.path.join(OUT, "libopesci_io_ref.so")
    cmd = ["g++", "-O2", "-fopenmp", "-fPIC", "-shared", "-std=c++11", "-include", "cmath", "-I", os.path.join(REF, "include"),
           os.path.join(REF, "src", "opesciIO.cpp"), os.path.join(REF, "src", "opesciHandy.cpp"),
           os.path.join(HERE, "io_ref_wrap.cpp"), "-o", lib]
    print("[make_io_ref] " + " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    sys.exit(0 if build() else 1)
