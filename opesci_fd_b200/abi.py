"""ctypes mirror of include/opesci_b200.h and the loader of the CUDA library.

Replaces the reference's "compile the generated file, then cdll.LoadLibrary it" step
(reference: opesci/grid.py:35-42, 99-103; opesci/compilation.py:26-48).  The product path
loads opesci_fd_b200/csrc/libopesci_b200.so (hand-written sm_100a CUDA behind a C ABI) and
fails loudly when it is missing -- there is no CPU fallback.
"""
import ctypes
import os
from ctypes import (POINTER, Structure, Union, c_char_p, c_double, c_float, c_int32, c_int64,
                    c_uint32, c_void_p)

OPESCI_MAX_M = 6
OPESCI_MAX_FIELDS = 9
OPESCI_MAX_TABLES = 12
OPESCI_MAX_PROG = 48

OP_TABLE, OP_CONST, OP_ADD, OP_SUB, OP_MUL, OP_NEG, OP_DIV, OP_FIELD = 1, 2, 3, 4, 5, 6, 7, 8
OP_MEDIA, OP_SQRT, OP_COS, OP_SIN, OP_ROUNDF = 9, 10, 11, 12, 13
MEDIA_NAMES = ['beta', 'lambda', 'mu', 'beta1', 'beta2', 'beta3', 'mu12', 'mu13', 'mu23']

KIND_STAGGERED_ELASTIC = 1
KIND_REGULAR_ACOUSTIC = 2
KIND_REGULAR_GENERIC = 3

ARITH_REFERENCE = 0
ARITH_FAST = 1
HOST_MIRROR_FULL = 0 << 4
HOST_MIRROR_NONE = 1 << 4
HOST_MIRROR_MASK = 0x30
NO_CUDA_GRAPH = 1 << 8
FORCE_UNFUSED = 1 << 9
OVERLAP = 1 << 10
FORCE_TILED = 1 << 11
L2_REFERENCE = 1 << 12
NO_ZFOLD = 1 << 13
NO_PAIR = 1 << 14

FS_NONE, FS_LEVANDER, FS_ROBERTSSON = 0, 1, 2


class OpesciGrid(Structure):
    _fields_ = [("field", c_void_p * OPESCI_MAX_FIELDS)]


class OpesciConvergence(Union):
    _fields_ = [("f32", c_float * OPESCI_MAX_FIELDS), ("f64", c_double * OPESCI_MAX_FIELDS)]


class OpesciProfiling(Structure):
    _fields_ = [("g_rtime", c_float), ("g_ptime", c_float), ("g_mflops", c_float)]


class OpesciSolInstr(Structure):
    _fields_ = [("op", c_int32), ("arg", c_int32), ("value", c_double)]


class OpesciSolProgram(Structure):
    _fields_ = [("n_instr", c_int32), ("n_tables", c_int32),
                ("table_axis", c_int32 * OPESCI_MAX_TABLES),
                ("table", POINTER(c_double) * OPESCI_MAX_TABLES),
                ("instr", OpesciSolInstr * OPESCI_MAX_PROG)]


class OpesciFieldSpec(Structure):
    _fields_ = [("lo", c_int32 * 3), ("hi", c_int32 * 3),
                ("l2_lo", c_int32 * 3), ("l2_hi", c_int32 * 3),
                ("init", OpesciSolProgram), ("final_", OpesciSolProgram)]


class OpesciB200Params(Structure):
    _fields_ = [
        ("struct_size", c_uint32), ("kind", c_int32), ("so", c_int32), ("is_double", c_int32),
        ("dim", c_int32 * 3), ("ntsteps", c_int32), ("nfields", c_int32), ("nlevels", c_int32),
        ("converge", c_int32), ("free_surface", c_int32), ("flags", c_int32),
        ("warmup_steps", c_int32), ("slab_rank", c_int32), ("slab_nranks", c_int32),
        ("dt", c_double), ("dx", c_double * 3), ("volume_literal", c_double),
        ("c_stress_normal", ((c_float * OPESCI_MAX_M) * 3) * 3),
        ("c_stress_shear", ((c_float * OPESCI_MAX_M) * 2) * 3),
        ("c_velocity", ((c_float * OPESCI_MAX_M) * 3) * 3),
        ("lev_stress", (((c_float * 2) * 3) * 3) * 3),
        ("lev_vnormal", (c_float * 3) * 3),
        ("lev_vtang", (c_float * 3) * 3),
        ("ac_coef", (c_float * OPESCI_MAX_M) * 3),
        ("ac_centre", c_float),
        ("ac_init_coef", (c_float * OPESCI_MAX_M) * 3),
        ("ac_init_centre", c_float),
        ("ac_init_const", c_double),
        ("hetero", c_int32), ("media_plane0", c_int32), ("media_nplanes", c_int32), ("fs_faces", c_int32),
        ("rho", POINTER(c_float)), ("vp", POINTER(c_float)), ("vs", POINTER(c_float)),
        ("h_c", (c_float * OPESCI_MAX_M) * 3),
        ("h_c2", (c_float * OPESCI_MAX_M) * 3),
        ("h_lev_den", (c_float * 2) * 3),
        ("h_lev_own", ((c_float * 2) * 3) * 3),
        ("h_lev_oth", ((c_float * 2) * 3) * 3),
        ("h_vn", (c_float * 2) * 3),
        ("n_receivers", c_int32), ("src_nt", c_int32), ("source_cell", c_int32 * 3), ("reserved2_", c_int32),
        ("receiver_cells", POINTER(c_int32)),
        ("src_x", POINTER(c_float)), ("src_y", POINTER(c_float)), ("src_z", POINTER(c_float)),
        ("receiver_out", c_void_p),
        ("fields", OpesciFieldSpec * OPESCI_MAX_FIELDS),
        ("generic_source", c_char_p),
    ]


EXPORTED_SYMBOLS = [
    "opesci_b200_configure", "opesci_execute", "opesci_convergence", "opesci_free",
    "opesci_b200_last_error", "opesci_b200_convergence_f64", "opesci_b200_last_timing",
    "opesci_b200_is_cuda", "opesci_b200_time_kernels",
    "opesci_b200_comm_unique_id", "opesci_b200_comm_init", "opesci_b200_comm_finalize",
    "opesci_b200_reserve_host", "opesci_b200_release_host", "opesci_b200_slab_range",
    "opesci_b200_execute_loopback", "opesci_b200_time_fused_parts", "opesci_b200_halo_transport",
]
# include/opesci_io.h (model input / field output around the path)
IO_SYMBOLS = [
    "opesci_b200_set_output", "opesci_b200_set_output_level", "opesci_b200_output_stats", "opesci_b200_dump_field_vts_3d",
    "opesci_b200_dump_field_vts",
    "opesci_b200_read_simple_binary_ptr", "opesci_b200_simple_binary_count", "opesci_b200_read_model_segy",
    "opesci_b200_segy_decode_device", "opesci_b200_ibm_to_float", "opesci_b200_read_xyz",
    "opesci_b200_resample_timeseries", "opesci_b200_calculate_dt", "opesci_b200_calculate_lame_constants",
]
SLAB_HALO = 8
COMM_ID_BYTES = 128
HALO_TRANSPORTS = {0: "none", 1: "nccl send/recv", 2: "peer memory (cudaIpc mapping, copy-engine pulls, NCCL tokens)", 3: "loopback copies"}

_PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CUDA_LIBRARY = os.path.join(_PKG_DIR, "csrc", "libopesci_b200.so")


def bind(lib):
    """Declare argument / result types on a loaded library exporting include/opesci_b200.h."""
    lib.opesci_b200_configure.argtypes = [POINTER(OpesciB200Params)]
    lib.opesci_b200_configure.restype = ctypes.c_int
    lib.opesci_execute.argtypes = [POINTER(OpesciGrid), POINTER(OpesciProfiling)]
    lib.opesci_execute.restype = ctypes.c_int
    lib.opesci_convergence.argtypes = [POINTER(OpesciGrid), POINTER(OpesciConvergence)]
    lib.opesci_convergence.restype = ctypes.c_int
    lib.opesci_free.argtypes = [POINTER(OpesciGrid)]
    lib.opesci_free.restype = ctypes.c_int
    lib.opesci_b200_last_error.argtypes = []
    lib.opesci_b200_last_error.restype = c_char_p
    lib.opesci_b200_convergence_f64.argtypes = [POINTER(OpesciGrid), POINTER(c_double)]
    lib.opesci_b200_convergence_f64.restype = ctypes.c_int
    lib.opesci_b200_last_timing.argtypes = [POINTER(c_double), POINTER(c_double), POINTER(c_int64)]
    lib.opesci_b200_last_timing.restype = ctypes.c_int
    if hasattr(lib, "opesci_b200_time_kernels"):
        lib.opesci_b200_time_kernels.argtypes = [POINTER(OpesciGrid), ctypes.c_int, POINTER(c_double)]
        lib.opesci_b200_time_kernels.restype = ctypes.c_int
    if hasattr(lib, "opesci_b200_time_fused_parts"):
        lib.opesci_b200_time_fused_parts.argtypes = [POINTER(OpesciGrid), ctypes.c_int, POINTER(c_double)]
        lib.opesci_b200_time_fused_parts.restype = ctypes.c_int
    if hasattr(lib, "opesci_b200_halo_transport"):
        lib.opesci_b200_halo_transport.argtypes = []
        lib.opesci_b200_halo_transport.restype = ctypes.c_int
    if hasattr(lib, "opesci_b200_reserve_host"):
        lib.opesci_b200_reserve_host.argtypes = [ctypes.c_size_t, ctypes.c_int]
        lib.opesci_b200_reserve_host.restype = ctypes.c_int
        lib.opesci_b200_release_host.argtypes = []
        lib.opesci_b200_release_host.restype = ctypes.c_int
    if hasattr(lib, "opesci_b200_comm_init"):
        lib.opesci_b200_comm_unique_id.argtypes = [c_void_p, ctypes.c_int]
        lib.opesci_b200_comm_unique_id.restype = ctypes.c_int
        lib.opesci_b200_comm_init.argtypes = [ctypes.c_int, ctypes.c_int, c_void_p, ctypes.c_int]
        lib.opesci_b200_comm_init.restype = ctypes.c_int
        lib.opesci_b200_comm_finalize.argtypes = []
        lib.opesci_b200_comm_finalize.restype = ctypes.c_int
    if hasattr(lib, "opesci_b200_slab_range"):
        lib.opesci_b200_slab_range.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int,
                                               POINTER(ctypes.c_int), POINTER(ctypes.c_int)]
        lib.opesci_b200_slab_range.restype = ctypes.c_int
    lib.opesci_b200_is_cuda.argtypes = []
    lib.opesci_b200_is_cuda.restype = ctypes.c_int
    if hasattr(lib, "opesci_b200_execute_loopback"):
        lib.opesci_b200_execute_loopback.argtypes = [ctypes.c_int, POINTER(OpesciGrid)]
        lib.opesci_b200_execute_loopback.restype = ctypes.c_int
    if hasattr(lib, "opesci_b200_set_output"):
        bind_io(lib)
    return lib


def bind_io(lib):
    """include/opesci_io.h"""
    c_int, c_float, c_size_t = ctypes.c_int, ctypes.c_float, ctypes.c_size_t
    PF, PI = POINTER(c_float), POINTER(c_int)
    lib.opesci_b200_set_output.argtypes = [c_char_p, c_int, c_int]
    lib.opesci_b200_set_output.restype = c_int
    lib.opesci_b200_set_output_level.argtypes = [c_int]
    lib.opesci_b200_set_output_level.restype = c_int
    lib.opesci_b200_output_stats.argtypes = [PI, PI]
    lib.opesci_b200_output_stats.restype = c_int
    lib.opesci_b200_dump_field_vts_3d.argtypes = [c_char_p, PI, PF, c_int, PF, c_int]
    lib.opesci_b200_dump_field_vts_3d.restype = c_int
    lib.opesci_b200_dump_field_vts.argtypes = [c_char_p, PI, PF, PF]
    lib.opesci_b200_dump_field_vts.restype = c_int
    lib.opesci_b200_read_simple_binary_ptr.argtypes = [c_char_p, PF, c_size_t]
    lib.opesci_b200_read_simple_binary_ptr.restype = c_int
    lib.opesci_b200_simple_binary_count.argtypes = [c_char_p]
    lib.opesci_b200_simple_binary_count.restype = c_int64
    lib.opesci_b200_read_model_segy.argtypes = [c_char_p, PF, c_size_t, PI, PF, c_int]
    lib.opesci_b200_read_model_segy.restype = c_int
    lib.opesci_b200_segy_decode_device.argtypes = [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_int, c_void_p]
    lib.opesci_b200_segy_decode_device.restype = c_int
    lib.opesci_b200_ibm_to_float.argtypes = [POINTER(ctypes.c_ubyte), c_int]
    lib.opesci_b200_ibm_to_float.restype = c_float
    lib.opesci_b200_read_xyz.argtypes = [c_char_p, PF, c_int]
    lib.opesci_b200_read_xyz.restype = c_int
    lib.opesci_b200_resample_timeseries.argtypes = [PF, c_int, c_float, c_double, PF, c_int]
    lib.opesci_b200_resample_timeseries.restype = c_int
    lib.opesci_b200_calculate_dt.argtypes = [PF, c_size_t, c_float]
    lib.opesci_b200_calculate_dt.restype = c_float
    lib.opesci_b200_calculate_lame_constants.argtypes = [PF, PF, PF, c_size_t, PF, PF]
    lib.opesci_b200_calculate_lame_constants.restype = None
    return lib


def load_library(path=None):
    """Load the CUDA library (or an explicitly given ABI-compatible one).

    Mirrors Grid._load_library (reference: opesci/grid.py:35-42): a load failure raises.
    """
    libname = path or CUDA_LIBRARY
    if "OPESCI_NCCL_LIB" not in os.environ:
        # slabs: the C library dlopens NCCL by name; point it at the copy shipped with the nvidia-nccl wheel
        # (the one torch uses) when there is one, found relative to the interpreter -- no fixed path
        import importlib.util
        spec = importlib.util.find_spec("nvidia")
        for base in (spec.submodule_search_locations if spec and spec.submodule_search_locations else []):
            cand = os.path.join(base, "nccl", "lib", "libnccl.so.2")
            if os.path.exists(cand):
                os.environ["OPESCI_NCCL_LIB"] = cand
                break
    if not os.path.exists(libname):
        raise Exception("Failed to load %s: file not found (build it with "
                        "`python -c 'import __graft_entry__ as g; g.build()'`); "
                        "there is no CPU fallback" % libname)
    try:
        lib = ctypes.CDLL(libname, mode=ctypes.RTLD_GLOBAL)
    except OSError as e:
        print("Library load error: ", e)
        raise Exception("Failed to load %s" % libname)
    return bind(lib)
