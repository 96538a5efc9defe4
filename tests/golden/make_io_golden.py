#!/usr/bin/env python
"""Golden vectors for include/opesci_io.h, produced by the REFERENCE's own libopesci helpers.

Runs in the development container only: needs oracle/_ref/libopesci_io_ref.so (oracle/refgen/make_io_ref.py,
the reference's src/opesciIO.cpp + src/opesciHandy.cpp compiled where they lie).  Writes
  tests/golden/io_model_be.segy, io_model_le.segy   synthetic SEG-Y model volumes (big / little endian, IBM floats)
  tests/golden/io_receivers.txt, io_sources.txt, io_src_{x,y,z}.bin
  tests/golden/io_golden.npz                          what the reference read / computed from them
Everything is deterministic (fixed seeds); floats are compared by bit pattern in tests/test_io.py.
"""
import ctypes
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.path.join(ROOT, "oracle", "_ref", "libopesci_io_ref.so")
FP = ctypes.POINTER(ctypes.c_float)


def fptr(a):
    return a.ctypes.data_as(FP)


def ibm_word(sign, exponent, mantissa):
    """IBM REAL*4 word: sign bit, 7-bit excess-64 base-16 exponent, 24-bit mantissa"""
    return (sign << 31) | ((exponent + 64) << 24) | mantissa


def write_segy(path, words, nx, ny, nz, big, scalar, x_origin, y_origin, step, dz):
    e = ">" if big else "<"
    head = bytearray(3600)
    struct.pack_into(e + "h", head, 3212, nx)
    struct.pack_into(e + "h", head, 3216, 1000)
    struct.pack_into(e + "h", head, 3220, nz)
    struct.pack_into(e + "h", head, 3224, 1)      # format code 1: IBM float
    with open(path, "wb") as f:
        f.write(head)
        for i in range(nx * ny):
            ix, iy = i % nx, i // nx
            th = bytearray(240)
            struct.pack_into(e + "h", th, 70, scalar)
            struct.pack_into(e + "i", th, 72, x_origin + ix * step)
            struct.pack_into(e + "i", th, 76, y_origin + iy * step)
            struct.pack_into(e + "h", th, 114, nz)
            struct.pack_into(e + "h", th, 116, dz)
            f.write(th)
            f.write(struct.pack(e + "%dI" % nz, *[int(w) for w in words[i]]))


def main():
    if not os.path.exists(REF):
        sys.exit("build oracle/_ref/libopesci_io_ref.so first (oracle/refgen/make_io_ref.py)")
    ref = ctypes.CDLL(REF)
    ref.ref_real2float.restype = ctypes.c_float
    ref.ref_calculate_dt.restype = ctypes.c_float
    ref.ref_calculate_dt.argtypes = [FP, ctypes.c_long, ctypes.c_float]
    ref.ref_resample.argtypes = [FP, ctypes.c_int, ctypes.c_float, ctypes.c_double, FP, ctypes.c_int]
    ref.ref_read_model_segy.argtypes = [ctypes.c_char_p, FP, ctypes.c_long, ctypes.POINTER(ctypes.c_int), FP]
    ref.ref_lame.argtypes = [FP, FP, FP, ctypes.c_long, FP, FP]
    rng = np.random.Generator(np.random.Philox(20261017))
    out = {}

    # ---- IBM words: random sign / exponent / mantissa plus the corners
    n = 4096
    words = (rng.integers(0, 2, n, dtype=np.uint64) << 31) | (rng.integers(0, 128, n, dtype=np.uint64) << 24) | rng.integers(0, 1 << 24, n, dtype=np.uint64)
    corners = [0, ibm_word(1, 0, 0), ibm_word(0, 1, 0x100000), ibm_word(1, 1, 0x100000), ibm_word(0, 63, 0xffffff), ibm_word(1, 63, 0xffffff),
               ibm_word(0, -64, 1), ibm_word(0, -64, 0xffffff), ibm_word(0, 32, 0xffffff), ibm_word(1, 33, 0x000001), ibm_word(0, -37, 0x800001),
               ibm_word(0, -38, 0x123456), ibm_word(1, -40, 0xfedcba), 0x41100000, 0xc276a000, 0x42640000]
    words = np.concatenate([np.array(corners, dtype=np.uint64), words]).astype(np.uint32)
    vals = np.empty(len(words), dtype=np.float32)
    for k, w in enumerate(words):
        vals[k] = ref.ref_real2float(struct.pack("<I", int(w)))     # the word as the reference's union sees it
    out["ibm_words"] = words
    out["ibm_values_bits"] = vals.view(np.uint32)

    # ---- SEG-Y model volumes
    for tag, big, nx, ny, nz, scalar, step, dz in (("be", True, 5, 4, 7, -10, 250, 20), ("le", False, 3, 6, 9, 2, 7, 3)):
        w = (rng.integers(0, 2, (nx * ny, nz), dtype=np.uint64) << 31) | (rng.integers(60, 70, (nx * ny, nz), dtype=np.uint64) << 24) | \
            rng.integers(0, 1 << 24, (nx * ny, nz), dtype=np.uint64)
        path = os.path.join(HERE, "io_model_%s.segy" % tag)
        write_segy(path, w, nx, ny, nz, big, scalar, 1000, -500, step, dz)
        dim = (ctypes.c_int * 3)()
        sp = np.zeros(3, dtype=np.float32)
        arr = np.zeros(nx * ny * nz, dtype=np.float32)
        rc = ref.ref_read_model_segy(path.encode(), fptr(arr), arr.size, dim, fptr(sp))
        assert rc == 0 and list(dim) == [nx, ny, nz], (rc, list(dim))
        out["segy_%s_dim" % tag] = np.array(list(dim), dtype=np.int32)
        out["segy_%s_spacing_bits" % tag] = sp.view(np.uint32)
        out["segy_%s_array_bits" % tag] = arr.view(np.uint32)

    # ---- resampling: a Ricker wavelet, longer / shorter / unchanged
    t = np.arange(48, dtype=np.float64) * 0.004
    a = (np.pi * 12.0 * (t - 0.08)) ** 2
    ricker = ((1 - 2 * a) * np.exp(-a)).astype(np.float32)
    out["resample_src"] = ricker
    cases = [(0.004, 0.001), (0.004, 0.0097), (0.004, 0.004), (0.0025, 0.004), (0.004, 0.0047), (0.004, 0.0036)]
    out["resample_cases"] = np.array(cases, dtype=np.float64)
    for k, (dt, sdt) in enumerate(cases):
        buf = np.zeros(1024, dtype=np.float32)
        n2 = ref.ref_resample(fptr(ricker), len(ricker), dt, sdt, fptr(buf), buf.size)
        assert n2 > 0
        out["resample_%d_bits" % k] = buf[:n2].copy().view(np.uint32)

    # ---- receiver / source files
    rec = os.path.join(HERE, "io_receivers.txt")
    with open(rec, "w") as f:
        f.write("x y z\n10.5 20.25 30\n\n40 50.125 60\n-1.5e2 2e-3 7\n")
    srcf = os.path.join(HERE, "io_sources.txt")
    with open(srcf, "w") as f:
        f.write("# header\n100 200 300.5\n")
    series = {}
    for c in "xyz":
        series[c] = rng.standard_normal(17).astype(np.float32)
        series[c].tofile(os.path.join(HERE, "io_src_%s.bin" % c))
    buf = np.zeros(30, dtype=np.float32)
    nrec = ref.ref_read_receivers(rec.encode(), fptr(buf), 10)
    out["receivers_bits"] = buf[:3 * nrec].copy().view(np.uint32)
    sx, sy, sz = (np.zeros(64, dtype=np.float32) for _ in range(3))
    ns = (ctypes.c_int * 3)()
    nsrc = ref.ref_read_sources(srcf.encode(), *[os.path.join(HERE, "io_src_%s.bin" % c).encode() for c in "xyz"], fptr(buf), 10,
                                fptr(sx), fptr(sy), fptr(sz), 64, ns)
    out["sources_xyz_bits"] = buf[:3 * nsrc].copy().view(np.uint32)
    out["sources_x_bits"] = sx[:ns[0]].copy().view(np.uint32)

    # ---- dt and Lame constants
    vp = (1.0 + rng.random(1000)).astype(np.float32) * 1500
    vs = (0.4 + 0.3 * rng.random(1000)).astype(np.float32) * 1500
    rho = (1.0 + 0.5 * rng.random(1000)).astype(np.float32) * 1000
    out["vp"], out["vs"], out["rho"] = vp, vs, rho
    out["dt_bits"] = np.array([ref.ref_calculate_dt(fptr(vp), vp.size, 12.5)], dtype=np.float32).view(np.uint32)
    mu, lam = np.zeros_like(vp), np.zeros_like(vp)
    ref.ref_lame(fptr(vp), fptr(vs), fptr(rho), vp.size, fptr(mu), fptr(lam))
    out["mu_bits"], out["lam_bits"] = mu.view(np.uint32), lam.view(np.uint32)

    np.savez_compressed(os.path.join(HERE, "io_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "io_golden.npz"), sorted(out))


if __name__ == "__main__":
    main()
