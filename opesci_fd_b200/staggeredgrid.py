"""StaggeredGrid: velocity-stress formulation of elastic waves on a staggered grid.

Mirrors the reference interface opesci/staggeredgrid.py:15-945 (constructor keywords,
`set_stress_fields`, `set_velocity_fields`, `calc_derivatives`, `solve_fd`,
`set_free_surface_boundary`, `set_media_params`, `get_time_step_limit`, the AI reports).
Where the reference derives update expressions symbolically and prints them as C++
(staggeredgrid.py:135-170, 612-945), this class reads the PDE coefficients off the equations
and lowers the model to the literal tables the generated code would have contained
(`build_params`), for the fixed-function sm_100a kernels:

  c_k * dt/dx_d * {lambda+2mu | lambda | mu | beta}     interior updates   (staggeredgrid.py:728-748)
  Levander / Robertsson free-surface parameters          ghost-cell loops   (staggeredgrid.py:750-864,
                                                                             fields.py:192-261, 294-381)
  analytic-solution programs + loop ranges               init / L2          (staggeredgrid.py:612-659, 892-945)
"""
from fractions import Fraction

import numpy as np
from sympy import Symbol, expand

from . import abi, cexpr
from .codeprinter import ccode, literal
from .derivative import DDerivative
from .fields import Media, SField, VField
from sympy import IndexedBase

from .regulargrid import RegularGrid, _frac, _same_field
from .util import staggered_first_weights

__all__ = ['StaggeredGrid']

# canonical slots (struct order of the reference driver, tests/eigenwave3d.py:54-67)
_NORMAL = [(1, 1), (2, 2), (3, 3)]
_SHEAR = [(1, 2), (2, 3), (1, 3)]


class StaggeredGrid(RegularGrid):
    _switches = ['omp', 'ivdep', 'simd', 'double', 'expand', 'eval_const',
                 'output_vts', 'converge', 'profiling', 'pluto', 'fission']
    _papi_events = []

    def __init__(self, stress_fields=None, velocity_fields=None, converge=False, **kwargs):
        self.sfields = []
        self.vfields = []
        self._free_surface = set()
        super(StaggeredGrid, self).__init__(**kwargs)
        self.converge = converge
        if stress_fields:
            self.set_stress_fields(stress_fields)
        if velocity_fields:
            self.set_velocity_fields(velocity_fields)

    @property
    def fields(self):
        """velocity fields first, then stress fields (reference: staggeredgrid.py:69-71)"""
        return self.vfields + self.sfields

    @property
    def io(self):
        return self.read or self.output_vts

    def set_alignment(self, alignment):
        self.alignment = alignment

    def get_time_step_limit(self):
        """reference: staggeredgrid.py:85-102"""
        if self.read:
            return 'physical parameters to be read from file, please compile and run the executable'
        l = self.defined_variable['lambda'].value
        m = self.defined_variable['mu'].value
        r = self.defined_variable['rho'].value
        Vp = ((l + 2 * m) / r) ** 0.5
        h = min([sp.value for sp in self.spacing])
        if self.order[1] == 2:
            return h / Vp / (3 ** 0.5)
        elif self.order[1] == 4:
            return 0.495 * h / Vp
        else:
            return 'not implemented yet'

    def set_stress_fields(self, sfields):
        """reference: staggeredgrid.py:104-114"""
        num = self.dimension + self.dimension * (self.dimension - 1) // 2
        if not len(sfields) == num:
            raise Exception('wrong number of stress fields: ' + str(num) + ' fields required.')
        self.sfields = sfields
        self.set_field_spacing()

    def set_velocity_fields(self, vfields):
        """reference: staggeredgrid.py:116-126"""
        num = self.dimension
        if not len(vfields) == num:
            raise Exception('wrong number of velocity fields: ' + str(num) + ' fields required.')
        self.vfields = vfields
        self.set_field_spacing()

    def calc_derivatives(self):
        """reference: staggeredgrid.py:128-133"""
        for field in self.sfields + self.vfields:
            field.populate_derivatives(max_order=1)

    def set_media_params(self, read=False, rho=1.0, vp=1.0, vs=0.5, rho_file='', vp_file='', vs_file=''):
        """reference: staggeredgrid.py:234-282"""
        self.read = read
        if self.read:
            # rho, vp, vs come from flat float32 files (opesci_read_simple_binary_ptr); the derived arrays
            # beta, beta1-3, lambda, mu, mu12/13/23 are computed by the library on the device
            self.rho_file, self.vp_file, self.vs_file = rho_file, vp_file, vs_file
            kw = dict(dimension=3, staggered=[False, False, False], index=self.index)
            self.rho, self.vp, self.vs = Media('rho', **kw), Media('vp', **kw), Media('vs', **kw)
            self.beta = [Media(n, **kw) for n in ('beta', 'beta1', 'beta2', 'beta3')]
            self.lam = Media('lambda', **kw)
            self.mu = [Media(n, **kw) for n in ('mu', 'mu12', 'mu13', 'mu23')]
            self.media_arrays = None      # optional (rho, vp, vs) numpy arrays instead of files
            self.media_plane0 = 0         # slab runs: first global x plane held by media_arrays
            return
        self.set_variable('rho', rho, 'float', True)
        self.set_variable('beta', 1.0 / rho, 'float', True)
        self.set_variable('lambda', rho * (vp ** 2 - 2 * vs ** 2), 'float', True)
        self.set_variable('mu', rho * (vs ** 2), 'float', True)

    # ------------------------------------------------------------------ PDE analysis
    def _slot_fields(self):
        """Check that the nine fields are the canonical U,V,W,Txx,Tyy,Tzz,Txy,Tyz,Txz."""
        if self.dimension != 3:
            raise NotImplementedError("B200 path: 3-D models only")
        vel = [f for f in self.vfields if isinstance(f, VField)]
        if [f.direction for f in vel] != [1, 2, 3]:
            raise NotImplementedError("velocity_fields must be given in direction order 1,2,3")
        dirs = [tuple(f.direction) for f in self.sfields if isinstance(f, SField)]
        if dirs != _NORMAL + _SHEAR:
            raise NotImplementedError("stress_fields must be Txx,Tyy,Tzz,Txy,Tyz,Txz (directions %s)"
                                      % (_NORMAL + _SHEAR,))
        normal = {d[0]: f for d, f in zip(dirs[:3], self.sfields[:3])}
        shear = {d: f for d, f in zip(dirs[3:], self.sfields[3:])}
        return {f.direction: f for f in vel}, normal, shear

    def solve_fd(self, equations):
        """reference: staggeredgrid.py:135-170.  Reads every PDE as
        `dF/dt = sum coef * dG/dx_d` and checks it has the velocity-stress structure the
        fixed-function kernels implement; keeps the coefficient EXPRESSIONS (evaluated at
        lowering time with the current media constants, like the reference's eval_const)."""
        if not len(self.fields) == len(equations):
            raise KeyError("Number of equations must be the same as number of fields.")
        self.eq = list(equations)
        vel, normal, shear = self._slot_fields()

        def stress_of(a, d):
            return normal[a] if a == d else shear[tuple(sorted((a, d)))]

        self.pde = {}
        for field, eq in zip(self.fields, self.eq):
            lhs = eq.lhs
            if not (isinstance(lhs, DDerivative) and _same_field(lhs.field, field) and lhs.axis == 0 and lhs.order == 1):
                raise NotImplementedError("equation %s: left side must be d%s/dt" % (eq, field.label))
            field.set_dt(eq.rhs)
            coefs = self._linear_coefficients(eq)
            table = {}
            for d, c in coefs.items():
                if d.order != 1 or d.axis == 0:
                    raise NotImplementedError("unsupported derivative %s" % d)
                table[(str(d.field.label), d.axis)] = c
            # expected sparsity
            if isinstance(field, VField):
                a = field.direction
                want = {(str(stress_of(a, d).label), d) for d in (1, 2, 3)}
            elif field.direction[0] == field.direction[1]:
                want = {(str(vel[d].label), d) for d in (1, 2, 3)}
            else:
                a, b = field.direction
                want = {(str(vel[a].label), b), (str(vel[b].label), a)}
            if set(table) != want:
                raise NotImplementedError(
                    "equation for %s is not of velocity-stress form (B200 kernels are fixed-function; "
                    "general PDEs are SURVEY.md 8f item 4)" % field.label)
            self.pde[str(field.label)] = table

    def set_media_arrays(self, rho, vp, vs, plane0=0):
        """B200 addition: hand the medium over as float32 arrays [nplanes][dim2][dim3] (same layout as
        the files) holding the global x planes [plane0, plane0+nplanes) -- a slab rank only needs its
        own planes (opesci_b200_slab_range)."""
        arrs = [np.ascontiguousarray(a, dtype=np.float32) for a in (rho, vp, vs)]
        self.media_arrays, self.media_plane0 = arrs, int(plane0)

    def _load_media(self):
        dims = [self.dim[d].value for d in range(3)]
        if self.media_arrays is None:
            n = dims[0] * dims[1] * dims[2]
            arrs = []
            for fn in (self.rho_file, self.vp_file, self.vs_file):
                if str(fn).lower().endswith(('.segy', '.sgy')):
                    a = self._read_segy(fn, dims)              # opesciIO.cpp:451-612: SEG-Y model volume
                else:
                    a = np.fromfile(fn, dtype='<f4', count=n)     # opesciIO.cpp:319: flat float32
                if a.size != n:
                    raise IOError("%s: expected %d float32 values, found %d" % (fn, n, a.size))
                arrs.append(np.ascontiguousarray(a.reshape(dims), dtype=np.float32))
            self.media_arrays, self.media_plane0 = arrs, 0
        for a in self.media_arrays:
            if a.ndim != 3 or list(a.shape[1:]) != dims[1:]:
                raise ValueError("media arrays must be [nplanes][dim2][dim3] = [*][%d][%d]" % (dims[1], dims[2]))
        return self.media_arrays

    def _read_segy(self, filename, dims):
        """SEG-Y model volume (IBM floats) through the library's reader (include/opesci_io.h), already in the
        [x][y][z] layout of rho / vp / vs; the volume must cover the whole array including the ghost planes,
        like the raw files the reference reads (vsize = dim1*dim2*dim3, staggeredgrid.py:549-551)."""
        import ctypes
        if self._library is None:
            self._load_library()
        lib = self._library
        dim = (ctypes.c_int * 3)()
        sp = (ctypes.c_float * 3)()
        if lib.opesci_b200_read_model_segy(str(filename).encode(), None, 0, dim, sp, 1) != 0:
            raise IOError("%s: not a readable SEG-Y model volume (format code 1)" % filename)
        if list(dim) != list(dims):
            raise IOError("%s: SEG-Y volume is %s, the grid needs %s" % (filename, list(dim), list(dims)))
        a = np.zeros(dims[0] * dims[1] * dims[2], dtype=np.float32)
        if lib.opesci_b200_read_model_segy(str(filename).encode(), a.ctypes.data_as(abi.POINTER(abi.c_float)), a.size, dim, sp, 1) != 0:
            raise IOError("%s: SEG-Y read failed" % filename)
        return a

    # ------------------------------------------------------------------ point source + receivers
    def _cell_of(self, coord):
        """grid node nearest to a coordinate, as an array index: round(coordinate/dx_d) + m
        (the reference's hand-written propagator uses round(coord/h), tests/src/test_ref_iso_elastic.cpp:229-231)"""
        m = self.margin.value
        return [int(round(float(c) / float(sp.value))) + m for c, sp in zip(coord, self.spacing)]

    def set_receivers(self, coordinates):
        """B200 addition: every receiver records U, V, W and (Txx+Tyy+Tzz)/3 at its nearest
        grid node at the end of every time step; read them with `receiver_data()` after `execute`/`run`."""
        self._receivers = [self._cell_of(c) for c in coordinates]

    def set_source(self, coordinate, xsrc, ysrc=None, zsrc=None):
        """Explosive point source: Txx, Tyy, Tzz[node] -= x|y|zsrc[ti]/3 at the end of step ti
        (tests/src/test_ref_iso_elastic.cpp:276-290)."""
        xs = np.ascontiguousarray(xsrc, dtype=np.float32)
        ys = xs if ysrc is None else np.ascontiguousarray(ysrc, dtype=np.float32)
        zs = xs if zsrc is None else np.ascontiguousarray(zsrc, dtype=np.float32)
        if not (xs.shape == ys.shape == zs.shape) or xs.ndim != 1:
            raise ValueError("source time series must be 1-D and of equal length")
        self._source = (self._cell_of(coordinate), xs, ys, zs)

    def receiver_data(self):
        """[ntsteps][4][n_receivers] array (U, V, W, mean normal stress) recorded by the last run."""
        return getattr(self, '_receiver_out', None)

    def _lower_hooks(self, p, keep):
        rec = getattr(self, '_receivers', None)
        src = getattr(self, '_source', None)
        self._receiver_out = None
        if rec:
            cells = np.ascontiguousarray(np.array(rec, dtype=np.int32).reshape(-1, 3))
            out = np.zeros((max(self.ntsteps.value, 1), 4, len(rec)), dtype=np.float64 if self.double else np.float32)
            p.n_receivers = len(rec)
            p.receiver_cells = cells.ctypes.data_as(abi.POINTER(abi.c_int32))
            p.receiver_out = out.ctypes.data_as(abi.c_void_p)
            self._receiver_out = out
            keep += [cells, out]
        if src:
            cell, xs, ys, zs = src
            p.src_nt = int(xs.shape[0])
            for d in range(3):
                p.source_cell[d] = cell[d]
            fptr = abi.POINTER(abi.c_float)
            p.src_x, p.src_y, p.src_z = xs.ctypes.data_as(fptr), ys.ctypes.data_as(fptr), zs.ctypes.data_as(fptr)
            keep += [xs, ys, zs]

    def set_free_surface_boundary(self, dimension, side):
        """reference: staggeredgrid.py:214-232.  Levander for so == 4, Robertsson otherwise."""
        self._free_surface.add((dimension, side))

    # ------------------------------------------------------------------ AI reports
    def get_stress_kernel_ai(self):
        """expanded form: 15*so ADD = MUL, 9 loads, 6 stores (regulargrid.py:293-327)"""
        return self._ai(15 * self.order[1], 9, 6)

    def get_velocity_kernel_ai(self):
        return self._ai(9 * self.order[1], 9, 3)

    def _ai(self, ops, load, store):
        word = 8 if self.double else 4
        ai = float(2 * ops) / (load + store) / word
        return (ai, ai, ops, ops, load, store)

    def get_overall_kernel_ai(self):
        """Interior kernels only, with the reference's ghost-cell adjustment
        (staggeredgrid.py:481-510); the boundary-loop weights (O(1/N)) are not included."""
        v, s = self.get_velocity_kernel_ai(), self.get_stress_kernel_ai()
        overall = (v[0] + s[0]) / 2.0
        adj = 1.0
        for d in self.dim[1:]:
            adj *= 1 - float(self.margin.value) / d.value
        return overall * adj, overall * adj

    # ------------------------------------------------------------------ lowering
    def build_params(self):
        if not getattr(self, 'pde', None):
            raise RuntimeError("solve_fd() must be called before the model can be lowered")
        for (d, s_) in self._free_surface:
            if d not in (1, 2, 3) or s_ not in (0, 1):
                raise ValueError("set_free_surface_boundary(dimension=1..3, side=0..1)")
        if len(set(self.order[1:])) != 1:
            raise NotImplementedError("equal spatial order on all axes required")
        if self.order[0] != 2:
            raise NotImplementedError("time order %d" % self.order[0])
        keep = []
        so = self.order[1]
        m = so // 2
        vel, normal, shear = self._slot_fields()
        p = self._common_params(abi.KIND_STAGGERED_ELASTIC, 9, 2)
        # any subset of the six faces (reference: one set_free_surface_boundary call per face, staggeredgrid.py:214-232;
        # faces without a call get no boundary loops, :766-768)
        p.fs_faces = sum(1 << (2 * (d - 1) + s_) for (d, s_) in self._free_surface)
        p.free_surface = abi.FS_NONE if not self._free_surface else (abi.FS_LEVANDER if so == 4 else abi.FS_ROBERTSSON)
        ck = staggered_first_weights(m)
        dt = _frac(self.dt.value)
        dx = [None] + [_frac(sp.value) for sp in self.spacing]

        def val(expr):
            return _frac(self._value(expr))

        def pde(field, operand, axis):
            return self.pde[str(field.label)][(str(operand.label), axis)]

        def fill(dst, coef, d):
            for k in range(m):
                dst[k] = literal(float(ck[k] * dt / dx[d] * coef))

        self._lower_hooks(p, keep)
        if self.read:
            self._lower_hetero(p, keep, vel, normal, shear, ck, dt, dx, m, so, pde)
        # interior updates
        M = {}
        for a in (1, 2, 3):
            if self.read:
                break
            for d in (1, 2, 3):
                M[(a, d)] = val(pde(normal[a], vel[d], d))
                fill(p.c_stress_normal[a - 1][d - 1], M[(a, d)], d)
                g = normal[a] if a == d else shear[tuple(sorted((a, d)))]
                fill(p.c_velocity[a - 1][d - 1], val(pde(vel[a], g, d)), d)
        for s, (a, b) in enumerate(_SHEAR):
            if self.read:
                break
            fill(p.c_stress_shear[s][0], val(pde(shear[(a, b)], vel[a], b)), b)
            fill(p.c_stress_shear[s][1], val(pde(shear[(a, b)], vel[b], a)), a)
        # Levander free surface (so == 4): eliminate d_d V_d with T_dd' = 0 (fields.py:313-353)
        # and build the ghost velocities from 2nd-order differences (fields.py:208-242)
        if so == 4 and not self.read:
            for d in (1, 2, 3):
                for e in (1, 2, 3):
                    if e == d:
                        continue
                    for f in (1, 2, 3):
                        if f == d:
                            continue
                        coef = M[(e, f)] - M[(e, d)] * M[(d, f)] / M[(d, d)]
                        fill(p.lev_stress[d - 1][e - 1][f - 1], coef, f)
                    ratio = float(dx[d] / dx[e])
                    p.lev_vnormal[d - 1][e - 1] = literal(float(M[(d, e)] / M[(d, d)] * dx[d] / dx[e]))
                    sh = shear[tuple(sorted((d, e)))]
                    c_tang = val(pde(sh, vel[d], e)) / val(pde(sh, vel[e], d))
                    p.lev_vtang[d - 1][e - 1] = literal(float(c_tang) * ratio)
        # init / L2: loops stop one short on staggered axes; coordinates are half-shifted there
        # (staggeredgrid.py:632-641, 918-927); first time dt/2 for velocities (staggeredgrid.py:647)
        loop = [Symbol('_' + x.name) for x in self.index]
        ti = self.ntsteps.value % 2
        dims = [self.dim[d].value for d in range(3)]
        for k, field in enumerate(self.fields):
            stag = [bool(field.staggered[d + 1]) for d in range(3)]
            lo = [m] * 3
            hi = [dims[d] - m - (1 if stag[d] else 0) for d in range(3)]
            coords = self._coordinates(stag)
            t0 = self.dt.value / 2 if field.staggered[0] else 0
            tn = self.dt.value * self.ntsteps.value if not field.staggered[0] \
                else self.dt.value * self.ntsteps.value + self.dt.value / 2.0
            ivars = self._solution_variables(coords)
            fvars = self._solution_variables(coords)
            fvars.field('__F__', cexpr.DOUBLE if self.double else cexpr.FLOAT)
            if self.read:
                for v in (ivars, fvars):
                    for mid, name in enumerate(abi.MEDIA_NAMES):
                        v.media(name, cexpr.FLOAT, mid)
                init_text = ccode(self._read_solution(field.sol.subs(self.t, t0), loop))
                placeholder = IndexedBase(str(field.label))[[ti] + loop]
                final_text = ccode(placeholder - self._read_solution(field.sol.subs(self.t, tn), loop))
                final_text = final_text.replace(ccode(placeholder), '__F__')
            else:
                init_text = ccode(field.sol.subs(self.t, t0))
                final_text = self._residual_text(field, ti, tn, loop)
            self._field_spec(p, k, field, lo, hi, lo, hi, init_text, ivars, final_text, fvars, keep)
        return p, keep

    def _read_solution(self, sol, loop):
        """`read` mode (reference: staggeredgrid.py:648-653): beta, lambda, mu become the per-cell arrays
        and the index symbols x,y,z become the INTEGER loop variables _x,_y,_z -- also where the
        solution meant them as coordinates."""
        sol = sol.subs({Symbol('beta'): self.beta[0][tuple(loop)], Symbol("lambda"): self.lam[tuple(loop)],
                        Symbol("mu"): self.mu[0][tuple(loop)]})
        for idx, l in zip(self.index, loop):
            sol = sol.subs(idx, l)
        return sol

    def _lower_hetero(self, p, keep, vel, normal, shear, ck, dt, dx, m, so, pde):
        """Heterogeneous (`read`) lowering: check that the PDE coefficients are exactly the isotropic
        elastic ones (lambda+2mu, lambda, mu, beta) and emit the media-free literals of
        `literal*G[...]*media[x][y][z]` plus the Levander tables."""
        if self.double:
            raise NotImplementedError("heterogeneous media: fp32 only (the reference reader is float*)")
        beta, lam, mu = Symbol('beta'), Symbol('lambda'), Symbol('mu')

        def same(a, b):
            return expand(a - b) == 0
        for a in (1, 2, 3):
            for d in (1, 2, 3):
                ok = same(pde(normal[a], vel[d], d), lam + 2 * mu if a == d else lam)
                g = normal[a] if a == d else shear[tuple(sorted((a, d)))]
                ok = ok and same(pde(vel[a], g, d), beta)
                if not ok:
                    raise NotImplementedError("heterogeneous media: isotropic elastic PDEs only")
        for (a, b) in _SHEAR:
            if not (same(pde(shear[(a, b)], vel[a], b), mu) and same(pde(shear[(a, b)], vel[b], a), mu)):
                raise NotImplementedError("heterogeneous media: isotropic elastic PDEs only")
        p.hetero = 1
        for d in (1, 2, 3):
            for k in range(m):
                p.h_c[d - 1][k] = literal(float(ck[k] * dt / dx[d]))
                p.h_c2[d - 1][k] = literal(float(2 * ck[k] * dt / dx[d]))
        P = {d: dx[1] * dx[2] * dx[3] / dx[d] for d in (1, 2, 3)}
        for d in (1, 2, 3):
            p.h_vn[d - 1][0] = literal(float(P[d]))
            p.h_vn[d - 1][1] = literal(float(2 * P[d]))
            if so != 4:
                continue
            p.h_lev_den[d - 1][0] = literal(float(12 * P[d]))
            p.h_lev_den[d - 1][1] = literal(float(24 * P[d]))
            for f in (1, 2, 3):
                for k in range(2):
                    p.h_lev_own[d - 1][f - 1][k] = literal(float(48 * P[d] * ck[k] * dt / dx[f]))
                    p.h_lev_oth[d - 1][f - 1][k] = literal(float(24 * P[d] * ck[k] * dt / dx[f]))
                if f != d:
                    p.lev_vtang[d - 1][f - 1] = literal(float(dx[d] / dx[f]))
        arrs = self._load_media()
        p.media_plane0, p.media_nplanes = self.media_plane0, arrs[0].shape[0]
        fptr = abi.POINTER(abi.c_float)
        p.rho, p.vp, p.vs = [a.ctypes.data_as(fptr) for a in arrs]
        keep += arrs
