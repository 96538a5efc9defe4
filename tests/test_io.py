"""include/opesci_io.h: model input / field output around the time-stepping path.

CPU part (`-m "not gpu"`): the host readers / resampler / VTS writer of libopesci_b200.so against the golden
vectors the REFERENCE's own libopesci produced (tests/golden/make_io_golden.py; bit patterns), and against
oracle/_ref/libopesci_io_ref.so itself where it exists (development container).
GPU part: SEG-Y decode kernel == host reader, per-step .vts snapshots == the oracle's field after that step.
"""
import ctypes
import os
import re
import struct
import zlib

import numpy as np
import pytest

from common import ROOT, make_grid
from opesci_fd_b200 import abi

GOLD = os.path.join(ROOT, "tests", "golden")
FP = ctypes.POINTER(ctypes.c_float)


def fptr(a):
    return a.ctypes.data_as(FP)


@pytest.fixture(scope="module")
def lib(cuda_lib):
    return cuda_lib


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLD, "io_golden.npz"))


def read_vts(path):
    """minimal reader for the appended-raw + zlib VTK XML files the library writes -> (extent, field, points)"""
    raw = open(path, "rb").read()
    marker = raw.index(b'<AppendedData encoding="raw">')
    xml = raw[:marker].decode()
    start = raw.index(b"_", marker) + 1
    ext = [int(v) for v in re.search(r'WholeExtent="([^"]+)"', xml).group(1).split()]
    assert 'compressor="vtkZLibDataCompressor"' in xml and 'byte_order="LittleEndian"' in xml
    offs = {m.group(1): int(m.group(2)) for m in re.finditer(r'Name="(\w+)"[^>]*offset="(\d+)"', xml)}

    def array(off):
        p = start + off
        nb, bs, last = struct.unpack_from("<3I", raw, p)
        sizes = struct.unpack_from("<%dI" % nb, raw, p + 12)
        p += 12 + 4 * nb
        out = []
        for k, c in enumerate(sizes):
            blk = zlib.decompress(raw[p:p + c])
            assert len(blk) == (last if (k == nb - 1 and last) else bs)
            out.append(blk)
            p += c
        return np.frombuffer(b"".join(out), dtype="<f4")
    return ext, array(offs["field"]), array(offs["Points"]).reshape(-1, 3)


# ------------------------------------------------------------------ CPU: readers vs the reference's golden vectors
def test_io_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, "include", "opesci_io.h")).read()
    declared = set(re.findall(r"\b(opesci_b200_\w+)\s*\(", hdr))
    assert declared == set(abi.IO_SYMBOLS)
    for sym in declared:
        assert hasattr(lib, sym), sym


def test_ibm_float_matches_reference(lib, gold):
    words = gold["ibm_words"]
    got = np.empty(len(words), dtype=np.float32)
    for k, w in enumerate(words):
        b = (ctypes.c_ubyte * 4)(*struct.pack("<I", int(w)))
        got[k] = lib.opesci_b200_ibm_to_float(b, 0)
        bs = (ctypes.c_ubyte * 4)(*struct.pack(">I", int(w)))
        assert np.float32(lib.opesci_b200_ibm_to_float(bs, 1)).view(np.uint32) == got[k].view(np.uint32)
    assert np.array_equal(got.view(np.uint32), gold["ibm_values_bits"])
    # known answers: 0x41100000 = 1.0, 0xC276A000 = -118.625, 0x42640000 = 100.0
    assert list(got[13:16]) == [1.0, -118.625, 100.0]


@pytest.mark.parametrize("tag", ["be", "le"])
def test_segy_model_reader_matches_reference(lib, gold, tag):
    path = os.path.join(GOLD, "io_model_%s.segy" % tag).encode()
    dim = (ctypes.c_int * 3)()
    sp = np.zeros(3, dtype=np.float32)
    assert lib.opesci_b200_read_model_segy(path, None, 0, dim, fptr(sp), 0) == 0      # sizes only
    nx, ny, nz = list(dim)
    assert [nx, ny, nz] == list(gold["segy_%s_dim" % tag])
    assert np.array_equal(sp.view(np.uint32), gold["segy_%s_spacing_bits" % tag])
    arr = np.zeros(nx * ny * nz, dtype=np.float32)
    assert lib.opesci_b200_read_model_segy(path, fptr(arr), arr.size - 1, dim, fptr(sp), 0) == -2
    assert lib.opesci_b200_read_model_segy(path, fptr(arr), arr.size, dim, fptr(sp), 0) == 0
    assert np.array_equal(arr.view(np.uint32), gold["segy_%s_array_bits" % tag])
    # layout 1 = [x][y][z], what rho / vp / vs of include/opesci_b200.h take
    xyz = np.zeros(nx * ny * nz, dtype=np.float32)
    assert lib.opesci_b200_read_model_segy(path, fptr(xyz), xyz.size, dim, fptr(sp), 1) == 0
    ref = arr.reshape(nz, ny, nx).transpose(2, 1, 0)
    assert np.array_equal(xyz.reshape(nx, ny, nz).view(np.uint32), np.ascontiguousarray(ref).view(np.uint32))


def test_segy_errors(lib, tmp_path):
    dim = (ctypes.c_int * 3)()
    sp = np.zeros(3, dtype=np.float32)
    assert lib.opesci_b200_read_model_segy(b"/nonexistent.segy", None, 0, dim, fptr(sp), 0) == -1
    bad = tmp_path / "bad.segy"
    head = bytearray(3600)
    struct.pack_into(">h", head, 3212, 2); struct.pack_into(">h", head, 3220, 2); struct.pack_into(">h", head, 3224, 5)   # IEEE: unsupported, like the reference
    bad.write_bytes(bytes(head) + bytes(2 * 248))
    assert lib.opesci_b200_read_model_segy(str(bad).encode(), None, 0, dim, fptr(sp), 0) == -1


def test_resample_matches_reference(lib, gold):
    src = np.ascontiguousarray(gold["resample_src"])
    for k, (dt, sdt) in enumerate(gold["resample_cases"]):
        want = gold["resample_%d_bits" % k]
        n2 = lib.opesci_b200_resample_timeseries(fptr(src), len(src), dt, sdt, None, 0)
        assert n2 == len(want)
        out = np.zeros(n2, dtype=np.float32)
        assert lib.opesci_b200_resample_timeseries(fptr(src), len(src), dt, sdt, fptr(out), n2 - 1) == -2 or n2 == len(src)
        assert lib.opesci_b200_resample_timeseries(fptr(src), len(src), dt, sdt, fptr(out), n2) == n2
        assert np.array_equal(out.view(np.uint32), want), (k, np.abs(out - want.view(np.float32)).max())


def test_xyz_and_binary_readers_match_reference(lib, gold):
    rec = os.path.join(GOLD, "io_receivers.txt").encode()
    n = lib.opesci_b200_read_xyz(rec, None, 0)
    assert n == 3
    buf = np.zeros(3 * n, dtype=np.float32)
    assert lib.opesci_b200_read_xyz(rec, fptr(buf), n - 1) == -2
    assert lib.opesci_b200_read_xyz(rec, fptr(buf), n) == n
    assert np.array_equal(buf.view(np.uint32), gold["receivers_bits"])
    src = os.path.join(GOLD, "io_sources.txt").encode()
    sb = np.zeros(3, dtype=np.float32)
    assert lib.opesci_b200_read_xyz(src, fptr(sb), 1) == 1
    assert np.array_equal(sb.view(np.uint32), gold["sources_xyz_bits"])
    fx = os.path.join(GOLD, "io_src_x.bin").encode()
    cnt = lib.opesci_b200_simple_binary_count(fx)
    assert cnt == len(gold["sources_x_bits"])
    sx = np.zeros(cnt, dtype=np.float32)
    assert lib.opesci_b200_read_simple_binary_ptr(fx, fptr(sx), cnt) == 0
    assert np.array_equal(sx.view(np.uint32), gold["sources_x_bits"])
    assert lib.opesci_b200_read_simple_binary_ptr(fx, fptr(sx), cnt + 1) == -2            # short file: refused, not over-read
    assert lib.opesci_b200_read_simple_binary_ptr(b"/nonexistent.bin", fptr(sx), 1) == -1
    assert lib.opesci_b200_read_xyz(b"/nonexistent.txt", None, 0) == -1


def test_dt_and_lame_match_reference(lib, gold):
    vp, vs, rho = (np.ascontiguousarray(gold[k]) for k in ("vp", "vs", "rho"))
    dt = np.float32(lib.opesci_b200_calculate_dt(fptr(vp), vp.size, 12.5))
    assert dt.view(np.uint32) == gold["dt_bits"][0]
    mu, lam = np.zeros_like(vp), np.zeros_like(vp)
    lib.opesci_b200_calculate_lame_constants(fptr(vp), fptr(vs), fptr(rho), vp.size, fptr(mu), fptr(lam))
    assert np.array_equal(mu.view(np.uint32), gold["mu_bits"])
    assert np.array_equal(lam.view(np.uint32), gold["lam_bits"])


def test_against_reference_library_when_present(lib):
    """development container only: random inputs through the reference's own code and through ours"""
    path = os.path.join(ROOT, "oracle", "_ref", "libopesci_io_ref.so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/libopesci_io_ref.so not built here")
    ref = ctypes.CDLL(path)
    ref.ref_resample.argtypes = [FP, ctypes.c_int, ctypes.c_float, ctypes.c_double, FP, ctypes.c_int]
    ref.ref_real2float.restype = ctypes.c_float
    rng = np.random.default_rng(7)
    for n, dt, sdt in ((31, 0.002, 0.0031), (64, 0.001, 0.00025), (50, 0.003, 0.0011), (9, 1.0, 3.0)):
        src = rng.standard_normal(n).astype(np.float32)
        a, b = np.zeros(4096, dtype=np.float32), np.zeros(4096, dtype=np.float32)
        na = ref.ref_resample(fptr(src), n, dt, sdt, fptr(a), a.size)
        nb = lib.opesci_b200_resample_timeseries(fptr(src), n, dt, sdt, fptr(b), b.size)
        assert na == nb and np.array_equal(a[:na].view(np.uint32), b[:nb].view(np.uint32))
    for w in rng.integers(0, 1 << 32, 20000, dtype=np.uint64):
        by = struct.pack("<I", int(w))
        assert np.float32(ref.ref_real2float(by)).view(np.uint32) == \
            np.float32(lib.opesci_b200_ibm_to_float((ctypes.c_ubyte * 4)(*by), 0)).view(np.uint32)


def test_vts_writer_round_trip(lib, tmp_path):
    dims = (ctypes.c_int * 3)(5, 4, 7)
    sp = np.array([0.5, 0.25, 2.0], dtype=np.float32)
    rng = np.random.default_rng(3)
    field = rng.standard_normal(5 * 4 * 7).astype(np.float32)
    name = str(tmp_path / "U_3")
    assert lib.opesci_b200_dump_field_vts_3d(name.encode(), dims, fptr(sp), 2, fptr(field), 0) == 0
    ext, f, pts = read_vts(name + ".vts")
    assert ext == [0, 6, 0, 3, 0, 4]                       # k (fastest) first, VTK's convention
    assert np.array_equal(f.view(np.uint32), field.view(np.uint32))
    # the reference's point loop (src/opesciIO.cpp:621-632): (i-margin)*spacing, k fastest
    i, j, k = np.meshgrid(np.arange(5), np.arange(4), np.arange(7), indexing="ij")
    want = np.stack([(i - 2).astype(np.float32) * sp[0], (j - 2).astype(np.float32) * sp[1], (k - 2).astype(np.float32) * sp[2]], -1)
    assert np.array_equal(pts, want.reshape(-1, 3))
    # a field larger than one compression block, with an x offset (slab piece)
    dims = (ctypes.c_int * 3)(3, 300, 301)
    big = rng.standard_normal(3 * 300 * 301).astype(np.float32)
    name = str(tmp_path / "U_big")
    assert lib.opesci_b200_dump_field_vts_3d(name.encode(), dims, fptr(sp), 2, fptr(big), 10) == 0
    ext, f, pts = read_vts(name + ".vts")
    assert np.array_equal(f, big) and pts[0, 0] == np.float32(8) * sp[0] and pts[-1, 2] == np.float32(298) * sp[2]
    assert lib.opesci_b200_dump_field_vts_3d(b"/nonexistent_dir/x", dims, fptr(sp), 2, fptr(big), 0) == -1
    # the reference's level 9 gives the same data in a smaller file
    size1 = os.path.getsize(name + ".vts")
    assert lib.opesci_b200_set_output_level(10) == -1 and lib.opesci_b200_set_output_level(9) == 0
    try:
        assert lib.opesci_b200_dump_field_vts_3d(name.encode(), dims, fptr(sp), 2, fptr(big), 10) == 0
    finally:
        lib.opesci_b200_set_output_level(1)
    _, f9, p9 = read_vts(name + ".vts")
    assert np.array_equal(f9, big) and np.array_equal(p9, pts) and os.path.getsize(name + ".vts") <= size1


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("tag,layout", [("be", 0), ("be", 1), ("le", 1)])
def test_gpu_segy_decode_matches_host_reader(lib, gold, tag, layout):
    import torch
    path = os.path.join(GOLD, "io_model_%s.segy" % tag)
    dim = (ctypes.c_int * 3)()
    sp = np.zeros(3, dtype=np.float32)
    assert lib.opesci_b200_read_model_segy(path.encode(), None, 0, dim, fptr(sp), 0) == 0
    nx, ny, nz = list(dim)
    host = np.zeros(nx * ny * nz, dtype=np.float32)
    assert lib.opesci_b200_read_model_segy(path.encode(), fptr(host), host.size, dim, fptr(sp), layout) == 0
    raw = np.fromfile(path, dtype=np.uint8)[3600:]
    d_raw = torch.from_numpy(raw).cuda()
    d_out = torch.zeros(nx * ny * nz, dtype=torch.float32, device="cuda")
    assert lib.opesci_b200_segy_decode_device(d_raw.data_ptr(), nx * ny, nx, nz, 1 if tag == "be" else 0, d_out.data_ptr(), layout, None) == 0
    torch.cuda.synchronize()
    assert np.array_equal(d_out.cpu().numpy().view(np.uint32), host.view(np.uint32))


@pytest.mark.gpu
def test_gpu_segy_decode_large_random(lib):
    """every exponent / sign, 2M samples: device decode == host decode bit for bit"""
    import torch
    rng = np.random.default_rng(11)
    nx, ny, nz = 64, 32, 1000
    rec = np.zeros((nx * ny, 240 + 4 * nz), dtype=np.uint8)
    rec[:, 240:] = rng.integers(0, 256, (nx * ny, 4 * nz), dtype=np.uint8)
    d_raw = torch.from_numpy(rec.reshape(-1)).cuda()
    d_out = torch.zeros(nx * ny * nz, dtype=torch.float32, device="cuda")
    assert lib.opesci_b200_segy_decode_device(d_raw.data_ptr(), nx * ny, nx, nz, 1, d_out.data_ptr(), 1, None) == 0
    torch.cuda.synchronize()
    words = rec[:, 240:].reshape(-1, 4)
    sub = rng.integers(0, len(words), 5000)
    got = d_out.cpu().numpy().reshape(nx, ny, nz)          # trace i -> (ix = i % nx, iy = i // nx); layout 1 is [x][y][z]
    for s in sub:
        i, iz = divmod(int(s), nz)
        want = np.float32(lib.opesci_b200_ibm_to_float((ctypes.c_ubyte * 4)(*words[s]), 1))
        assert got[i % nx, i // nx, iz].view(np.uint32) == want.view(np.uint32)


@pytest.mark.gpu
@pytest.mark.parametrize("kind,so,double", [("eigenwave3d", 4, False), ("eigenwave3d", 8, False), ("eigenwave3d", 4, True),
                                            ("simplewave3d", 4, False)])
def test_gpu_per_step_vts_output_matches_oracle(lib, oracle_lib, tmp_path, kind, so, double):
    """output_vts switch (regulargrid.py:702-719): file <ti> holds the first field's new level after step ti"""
    steps, n = 5, 20
    cfg = dict(kind=kind, so=so, grid_size=[n, n + 2, n + 5], dt=0.25 / 64, steps=steps, double=double, domain=[1.0, 1.0, 1.0])
    prefix = str(tmp_path / "U_")
    assert lib.opesci_b200_set_output(prefix.encode(), 0, 1) == 0
    try:
        g = make_grid(cfg, flags=abi.ARITH_REFERENCE)
        g.run(library=lib)
    finally:
        lib.opesci_b200_set_output(None, 0, 0)
    nf, ne = ctypes.c_int(), ctypes.c_int()
    lib.opesci_b200_output_stats(ctypes.byref(nf), ctypes.byref(ne))
    assert (nf.value, ne.value) == (steps, 0)
    final = g.field_array(0)
    g.free()
    m = so // 2
    dims = [c + 1 + 2 * m for c in cfg["grid_size"]]
    period = 2 if kind == "eigenwave3d" else 3
    for ti in range(steps):
        ext, f, pts = read_vts("%s%d.vts" % (prefix, ti))
        assert ext == [0, dims[2] - 1, 0, dims[1] - 1, 0, dims[0] - 1]
        # oracle: the same model stopped after ti+1 steps
        c2 = dict(cfg, steps=ti + 1)
        o = make_grid(c2, flags=abi.ARITH_REFERENCE)
        o.run(library=oracle_lib)
        lvl = (ti + 1) % 2 if period == 2 else (ti + 2) % 3
        want = o.field_array(0)[lvl].astype(np.float32)
        o.free()
        assert np.array_equal(f.reshape(dims).view(np.uint32), want.view(np.uint32)), ti
    # the last snapshot is also the level opesci_execute hands back
    lvl = steps % 2 if period == 2 else (steps + 1) % 3
    assert np.array_equal(f.reshape(dims), final[lvl].astype(np.float32))
    # margin is the literal 2 of the emitted call, whatever the order (regulargrid.py:718)
    assert pts[0, 0] == np.float32(-2) * np.float32(1.0 / cfg["grid_size"][0])


@pytest.mark.gpu
def test_gpu_vts_output_every_k_and_disarm(lib, tmp_path):
    cfg = dict(kind="eigenwave3d", so=4, grid_size=[24, 24, 24], dt=0.25 / 64, steps=7, double=False, domain=[1.0, 1.0, 1.0])
    prefix = str(tmp_path / "snap_")
    lib.opesci_b200_set_output(prefix.encode(), 2, 3)          # field W every 3rd step
    g = make_grid(cfg, flags=abi.ARITH_FAST | abi.HOST_MIRROR_NONE)
    g.run(library=lib)
    g.free()
    lib.opesci_b200_set_output(None, 0, 0)
    assert sorted(os.listdir(tmp_path)) == ["snap_0.vts", "snap_3.vts", "snap_6.vts"]
    g = make_grid(cfg, flags=abi.ARITH_FAST | abi.HOST_MIRROR_NONE)
    g.run(library=lib)
    g.free()
    assert len(os.listdir(tmp_path)) == 3                       # disarmed: nothing new


@pytest.mark.gpu
def test_gpu_output_vts_switch_through_the_front_end(lib, tmp_path):
    """grid.set_switches(output_vts=True) -> "U_<ti>.vts" per step, like the generated code (staggeredgrid.py:882-890)"""
    cfg = dict(kind="eigenwave3d", so=4, grid_size=[16, 16, 16], dt=0.25 / 64, steps=3, double=False, domain=[1.0, 1.0, 1.0])
    g = make_grid(cfg, flags=abi.ARITH_REFERENCE)
    g.set_switches(output_vts=True)
    g.output_prefix = str(tmp_path) + os.sep
    g.run(library=lib)
    last = g.field_array(0)[3 % 2].copy()
    g.free()
    assert sorted(os.listdir(tmp_path)) == ["U_0.vts", "U_1.vts", "U_2.vts"]
    _, f, _ = read_vts(str(tmp_path / "U_2.vts"))
    assert np.array_equal(f.view(np.uint32), last.reshape(-1).view(np.uint32))


@pytest.mark.gpu
def test_gpu_heterogeneous_media_from_segy(lib, tmp_path):
    """`read` mode fed from SEG-Y model volumes == the same medium handed over as arrays, bit for bit"""
    import sys
    sys.path.insert(0, GOLD)
    from make_io_golden import write_segy
    import eigenwave3d as drv
    n, so = 12, 4
    dims = [n + 1 + so] * 3
    rng = np.random.default_rng(5)
    media, files = [], []
    for name, e, lo, hi in (("rho", 65, 1 << 20, 3 << 19), ("vp", 65, 1 << 20, 3 << 19), ("vs", 64, int(0.4 * 2 ** 24), int(0.7 * 2 ** 24))):
        words = (np.uint64(e) << np.uint64(24)) | rng.integers(lo, hi, (dims[0] * dims[1], dims[2]), dtype=np.uint64)
        path = str(tmp_path / (name + ".segy"))
        write_segy(path, words, dims[0], dims[1], dims[2], True, 1, 0, 0, 10, 10)
        files.append(path)
        vals = np.array([lib.opesci_b200_ibm_to_float((ctypes.c_ubyte * 4)(*struct.pack(">I", int(w))), 1) for w in words.reshape(-1)],
                        dtype=np.float32).reshape(dims[1], dims[0], dims[2])          # trace i = ix + iy*nx
        media.append(np.ascontiguousarray(vals.transpose(1, 0, 2)))
    assert 1.0 <= media[0].min() and media[0].max() < 1.5 and 0.4 <= media[2].min() and media[2].max() < 0.7
    out = []
    for use_files in (False, True):
        g = drv.eigenwave3d((1.0, 1.0, 1.0), (n, n, n), 0.004, 0.04, accuracy_order=[2, so, so, so], o_converge=False, read=True,
                            rho_file=files[0], vp_file=files[1], vs_file=files[2], verbose=False)
        g.ntsteps.value = 10
        g.b200_flags = abi.ARITH_REFERENCE
        if not use_files:
            g.set_media_arrays(*media)
        g.run(library=lib)
        out.append(np.stack([g.field_array(k).copy() for k in range(9)]))
        g.free()
    assert np.isfinite(out[0]).all() and np.abs(out[0]).max() > 0
    assert np.array_equal(out[0].view(np.uint32), out[1].view(np.uint32))


def test_segy2vts_converter(lib, gold, tmp_path):
    """src/segy2vts.cpp: SEG-Y volume -> <base>.vts through the reader and the x-fastest writer"""
    import shutil
    from opesci_fd_b200 import segy2vts
    src = str(tmp_path / "model.segy")
    shutil.copy(os.path.join(GOLD, "io_model_be.segy"), src)
    out, dim, spacing = segy2vts.convert(src, library=lib)
    assert out == str(tmp_path / "model.vts") and dim == list(gold["segy_be_dim"])
    ext, f, pts = read_vts(out)
    assert ext == [0, dim[0] - 1, 0, dim[1] - 1, 0, dim[2] - 1]          # x fastest: VTK's own convention
    assert np.array_equal(f.view(np.uint32), gold["segy_be_array_bits"])
    k, j, i = np.meshgrid(np.arange(dim[2]), np.arange(dim[1]), np.arange(dim[0]), indexing="ij")
    sp = np.array(spacing, dtype=np.float32)
    want = np.stack([i.astype(np.float32) * sp[0], j.astype(np.float32) * sp[1], k.astype(np.float32) * sp[2]], -1).reshape(-1, 3)
    assert np.array_equal(pts, want)
    with pytest.raises(ValueError):
        segy2vts.convert(str(tmp_path / "model.dat"), library=lib)
