"""SEG-Y model volume -> VTK structured grid, `python -m opesci_fd_b200.segy2vts model.segy`.

Mirrors the reference's converter (src/segy2vts.cpp): the base name is the file name up to its
.sgy / .segy / .SGY / .SEGY extension, the volume is read with the SEG-Y reader
(opesci_read_model_segy -> opesci_b200_read_model_segy, layout 0) and written with the x-fastest
writer (opesci_dump_field_vts -> opesci_b200_dump_field_vts) to <base>.vts.
"""
import ctypes
import sys

import numpy as np

from . import abi


def convert(filename, library=None):
    base = None
    for ext in (".sgy", ".segy", ".SGY", ".SEGY"):          # same search order as segy2vts.cpp:50-66
        pos = filename.rfind(ext)
        if pos >= 0:
            base = filename[:pos]
            break
    if base is None:
        raise ValueError("Do not recognise file extension. Expecting either .segy or .sgy")
    lib = library or abi.load_library()
    dim = (ctypes.c_int * 3)(1, 1, 1)
    spacing = (ctypes.c_float * 3)(1.0, 1.0, 1.0)
    if lib.opesci_b200_read_model_segy(filename.encode(), None, 0, dim, spacing, 0) != 0:
        raise IOError("%s: not a readable SEG-Y model volume" % filename)
    array = np.zeros(dim[0] * dim[1] * dim[2], dtype=np.float32)
    fptr = array.ctypes.data_as(ctypes.POINTER(ctypes.c_float))
    if lib.opesci_b200_read_model_segy(filename.encode(), fptr, array.size, dim, spacing, 0) != 0:
        raise IOError("%s: SEG-Y read failed" % filename)
    if lib.opesci_b200_dump_field_vts(base.encode(), dim, spacing, fptr) != 0:
        raise IOError("cannot write %s.vts" % base)
    return base + ".vts", list(dim), list(spacing)


if __name__ == "__main__":
    if len(sys.argv) != 2:
        sys.exit("usage: python -m opesci_fd_b200.segy2vts model.segy")
    out, dim, spacing = convert(sys.argv[1])
    print("%s: %d x %d x %d, spacing %g %g %g" % (out, dim[0], dim[1], dim[2], spacing[0], spacing[1], spacing[2]))
