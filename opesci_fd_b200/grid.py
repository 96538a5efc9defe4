"""Execution infrastructure shared by RegularGrid and StaggeredGrid.

Mirrors the reference interface opesci/grid.py:8-156 (`compiler` property, `generate`,
`compile`, `execute`, `convergence`) with the same call sequence and printed output.  The
difference is what sits behind it: instead of writing C++ and compiling it, the model is
lowered to an `OpesciB200Params` block (include/opesci_b200.h) and handed to the prebuilt CUDA
library through `opesci_b200_configure`; `opesci_execute` / `opesci_convergence` are then
called exactly like the reference calls its generated functions (grid.py:118-122, 150-156).
"""
import ctypes
import os
import json
from ctypes import byref
from os import environ

from . import abi
from .compilation import B200Compiler, ClangCompiler, GNUCompiler, IntelCompiler

__all__ = ['Grid']


class Grid(object):
    _compiler = GNUCompiler()

    src_code = None
    src_file = None
    src_lib = None

    _library = None
    _arg_grid = None
    _arg_conv = None
    _params = None
    _params_keepalive = None

    # flags forwarded to the library (include/opesci_b200.h)
    b200_flags = abi.ARITH_REFERENCE | abi.HOST_MIRROR_FULL

    def _load_library(self, src_lib=None):
        """reference: opesci/grid.py:35-42"""
        self._library = abi.load_library(src_lib or self.src_lib)

    @property
    def compiler(self):
        return self._compiler

    @compiler.setter
    def compiler(self, compiler):
        if compiler in ['g++', 'gnu']:
            self._compiler = GNUCompiler()
        elif compiler in ['icpc', 'intel']:
            self._compiler = IntelCompiler()
        elif compiler in ['clang', 'clang++']:
            self._compiler = ClangCompiler()
        elif compiler in ['nvcc', 'b200', 'cuda']:
            self._compiler = B200Compiler()
        else:
            raise ValueError("Unknown compiler.")

    # ------------------------------------------------------------------ lowering
    def build_params(self):
        """-> (OpesciB200Params, keepalive objects).  Implemented by the grid classes."""
        raise NotImplementedError

    def describe(self):
        """JSON-able summary of the lowered model (what `generate` writes)."""
        p, _ = self.build_params()
        m = p.so // 2

        def arr(a, *shape):
            if not shape:
                return float(a)
            return [arr(a[i], *shape[1:]) for i in range(shape[0])]
        out = dict(kind=int(p.kind), so=int(p.so), is_double=int(p.is_double), dim=list(p.dim),
                   ntsteps=int(p.ntsteps), nfields=int(p.nfields), nlevels=int(p.nlevels),
                   dt=p.dt, dx=list(p.dx), free_surface=int(p.free_surface),
                   volume_literal=p.volume_literal)
        if p.kind == abi.KIND_STAGGERED_ELASTIC:
            out.update(c_stress_normal=arr(p.c_stress_normal, 3, 3, m),
                       c_stress_shear=arr(p.c_stress_shear, 3, 2, m),
                       c_velocity=arr(p.c_velocity, 3, 3, m),
                       lev_stress=arr(p.lev_stress, 3, 3, 3, 2),
                       lev_vnormal=arr(p.lev_vnormal, 3, 3), lev_vtang=arr(p.lev_vtang, 3, 3))
            if p.hetero:
                out.update(hetero=1, media_files=[self.rho_file, self.vp_file, self.vs_file],
                           h_c=arr(p.h_c, 3, m), h_c2=arr(p.h_c2, 3, m), h_lev_den=arr(p.h_lev_den, 3, 2),
                           h_lev_own=arr(p.h_lev_own, 3, 3, 2), h_lev_oth=arr(p.h_lev_oth, 3, 3, 2),
                           h_vn=arr(p.h_vn, 3, 2))
        else:
            out.update(ac_coef=arr(p.ac_coef, 3, m), ac_centre=float(p.ac_centre),
                       ac_init_coef=arr(p.ac_init_coef, 3, m), ac_init_centre=float(p.ac_init_centre),
                       ac_init_const=p.ac_init_const)
        return out

    # ------------------------------------------------------------------ reference API
    def generate(self, filename, compiler=None):
        """reference: opesci/grid.py:59-69.  Writes the lowered model (JSON) instead of C++."""
        if compiler:
            self.compiler = compiler
        self.src_code = json.dumps(self.describe(), indent=1, sort_keys=True)
        self.src_file = filename
        with open(self.src_file, 'w') as f:
            f.write(self.src_code)
        print("Generated:", self.src_file)

    def compile(self, filename, compiler=None, shared=True):
        """reference: opesci/grid.py:71-83"""
        if compiler:
            self.compiler = compiler
        if self.src_file is None:
            self.generate(filename)
        out = self.compiler.compile(self.src_file, shared=shared)
        if shared:
            self.src_lib = out
        return out

    def execute(self, filename, compiler='g++', nthreads=1, affinity='close'):
        """reference: opesci/grid.py:85-130"""
        environ["OMP_NUM_THREADS"] = str(nthreads)
        if affinity in ['close', 'spread']:
            environ["OMP_PROC_BIND"] = affinity
        elif affinity in ['compact', 'scatter']:
            environ["KMP_AFFINITY"] = "granularity=thread,%s" % affinity
        else:
            print("""ERROR: Only the following affinity settings are supported:
 * OMP_PROC_BIND: 'close', 'spread'
 * KMP_AFFINITY: 'compact', 'scatter'""")
            raise ValueError("Unknown thread affinity setting: %s")

        if self.src_lib is None:
            self.compile(filename, compiler=compiler, shared=True)
        self._load_library(src_lib=self.src_lib)
        if not self._library.opesci_b200_is_cuda():
            # execute() is the product path: only the CUDA library may serve it (no CPU fallback)
            raise RuntimeError("%s is not the CUDA library; Grid.execute() runs on the B200 only" % self.src_lib)
        print("Executing on B200 (CUDA) (host threads requested: %d, affinity=%s)" % (nthreads, affinity))
        self.run()
        if self.profiling:
            print("B200:: time loop: %f (sec)" % self._arg_profiling.g_rtime)
            print("B200:: opesci_execute: %f (sec)" % self._arg_profiling.g_ptime)
            print("B200:: Total MFlops/s: %f" % self._arg_profiling.g_mflops)

    def run(self, library=None):
        """configure + opesci_execute on an already loaded (or given) library."""
        if library is not None:
            self._library = library
        if self._library is None:
            self._load_library()
        if self._arg_grid is not None:
            self.free()
        self._params, self._params_keepalive = self.build_params()
        self._params.flags = int(self.b200_flags)
        if os.environ.get("OPESCI_L2_REFERENCE", "0") not in ("", "0"):
            # convergence() prints the reference's own digits (serial real_t accumulation, staggeredgrid.py:916,935)
            self._params.flags |= abi.L2_REFERENCE
        lib = self._library
        if lib.opesci_b200_configure(byref(self._params)) != 0:
            raise RuntimeError("opesci_b200_configure: %s" % lib.opesci_b200_last_error().decode())
        self._arg_grid = abi.OpesciGrid()
        self._arg_profiling = abi.OpesciProfiling()
        # `output_vts` switch: the generator emits a per-step dump of the first field into "<label>_<ti>.vts"
        # (reference: opesci/regulargrid.py:702-719, staggeredgrid.py:882-890); here the library streams it
        # out asynchronously (include/opesci_io.h)
        armed = bool(getattr(self, "output_vts", False)) and hasattr(lib, "opesci_b200_set_output")
        if armed:
            prefix = "%s%s_" % (getattr(self, "output_prefix", ""), str(self.fields[0].label))
            lib.opesci_b200_set_output(prefix.encode(), 0, int(getattr(self, "output_every", 1)))
        try:
            rc = lib.opesci_execute(byref(self._arg_grid), byref(self._arg_profiling))
        finally:
            if armed:
                lib.opesci_b200_set_output(None, 0, 0)
        if rc != 0:
            self._arg_grid = None
            raise RuntimeError("opesci_execute: %s" % lib.opesci_b200_last_error().decode())
        return self._arg_grid

    def convergence(self):
        """reference: opesci/grid.py:132-156"""
        if self._library is None:
            self._load_library()
        if self._arg_grid is None:
            raise RuntimeError("""Convergence could not find grid argument!
You need to you run grid.execute() first!""")
        arg_conv = abi.OpesciConvergence()
        print("Convergence:")
        if self._library.opesci_convergence(byref(self._arg_grid), byref(arg_conv)) != 0:
            raise RuntimeError("opesci_convergence: %s" % self._library.opesci_b200_last_error().decode())
        values = arg_conv.f64 if self.double else arg_conv.f32
        result = {}
        for k, f in enumerate(self.fields):
            name = '%s_l2' % str(f.label)
            result[name] = float(values[k])
            print("%s: %.10f" % (name, values[k]))
        self._arg_conv = arg_conv
        return result

    def convergence_f64(self):
        out = (ctypes.c_double * abi.OPESCI_MAX_FIELDS)()
        if self._library.opesci_b200_convergence_f64(byref(self._arg_grid), out) != 0:
            raise RuntimeError(self._library.opesci_b200_last_error().decode())
        return [float(out[k]) for k in range(len(self.fields))]

    def free(self):
        """opesci_free (the reference's python side never calls it, grid.py:85-156)."""
        if self._arg_grid is not None and self._library is not None:
            self._library.opesci_free(byref(self._arg_grid))
        self._arg_grid = None

    def field_array(self, k):
        """numpy view [nlevels][dim1][dim2][dim3] of host field k (HOST_MIRROR_FULL only).

        The view aliases the library's result array: it is valid until free() / the next run() of this grid
        (opesci_free hands the block back to the library's host pool).  Copy it to keep it longer.
        """
        import numpy as np
        if self._arg_grid is None:
            raise RuntimeError("field_array: no results (run() first; free() releases them)")
        p = self._params
        if (int(p.flags) & abi.HOST_MIRROR_MASK) != abi.HOST_MIRROR_FULL:
            raise RuntimeError("field_array: the fields were left on the device (HOST_MIRROR_NONE); "
                               "grid->field[] are device pointers")
        n = p.nlevels * p.dim[0] * p.dim[1] * p.dim[2]
        ctype = ctypes.c_double if p.is_double else ctypes.c_float
        buf = ctypes.cast(self._arg_grid.field[k], ctypes.POINTER(ctype * n)).contents
        return np.frombuffer(buf, dtype=np.float64 if p.is_double else np.float32).reshape(
            p.nlevels, p.dim[0], p.dim[1], p.dim[2])
