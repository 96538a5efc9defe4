"""RegularGrid: unstaggered grid with one scalar field and a 2nd-order-in-time wave equation.

Mirrors the reference interface opesci/regulargrid.py:16-719: same constructor keywords,
switches, `set_*` methods, `calc_derivatives`, `solve_fd`, `get_time_step_limit`,
`get_kernel_ai`, and the attributes drivers read (`dim`, `spacing`, `margin`, `dt`, `ntsteps`,
`order`, `time`, `tp`, `real_t`, `defined_variable`).  Instead of emitting C++ text
(regulargrid.py:391-719) it lowers the model to the parameter block of the CUDA library
(`build_params`): dims, the float literals of the update stencil, the second-initialisation
stencil, and the analytic-solution programs.
"""
import mmap
from fractions import Fraction

import numpy as np
from sympy import Indexed, IndexedBase, Symbol, expand, symbols

from . import abi, cexpr
from .codeprinter import ccode, literal
from .derivative import DDerivative
from .fields import RegularField
from .grid import Grid
from .util import central_weights, get_all_objects, variable_to_symbol
from .variable import Variable

__all__ = ['RegularGrid']


def _frac(x):
    return Fraction(float(x))


def _same_field(a, b):
    """sympy may hand back an equal-by-name symbol created earlier in the process (its caches
    key on structural equality), so fields are compared by label, never by identity."""
    return a is not None and b is not None and str(a.label) == str(b.label)


class RegularGrid(Grid):
    _papi_events = []
    _switches = ['omp', 'ivdep', 'simd', 'double', 'expand', 'eval_const',
                 'output_vts', 'converge', 'profiling', 'pluto', 'fission']
    _params = ['c', 'v']

    def __init__(self, dimension, index=None, fields=None, double=False, profiling=False, pluto=False,
                 fission=False, omp=True, ivdep=True, simd=False, io=False, expand=True, eval_const=True,
                 grid_size=(10, 10, 10), domain_size=None, output_vts=False):
        super(RegularGrid, self).__init__()
        self.dimension = dimension
        self.double = double
        self.real_t = 'double' if self.double else 'float'
        # NB: dt keeps the type it gets HERE even if the `double` switch is set later
        # (reference: regulargrid.py:29 -- `const float dt` in double-mode generated code)
        self.dt = Variable('dt', 0.01, self.real_t, True)
        self.ntsteps = Variable('ntsteps', 100, 'int', True)
        self.alignment = mmap.PAGESIZE
        self.margin = Variable('margin', 2, 'int', True)
        self.dim = [Variable('dim' + str(k + 1), 10 + 1 + self.margin.value * 2, 'int', True)
                    for k in range(self.dimension)]
        self.size = [1.0] * dimension
        default_order = [2] + [4] * self.dimension
        self.t = Symbol('_t')
        self.grid_size = grid_size
        self.max_derivative_order = 1
        if fields is not None:
            self.fields = fields
        self.set_order(default_order)
        self.set_grid_size(grid_size)
        self.set_field_spacing()
        self.set_index(index)
        self.defined_variable = {}
        self.pluto = pluto
        self.omp = omp
        self.ivdep = ivdep
        self.simd = simd
        self.output_vts = output_vts
        self.expand = expand
        self.eval_const = eval_const
        self.profiling = profiling
        self.fission = fission
        self.converge = False
        if domain_size:
            self.set_domain_size(domain_size)
        self.read = False
        self.eq = []

    # ------------------------------------------------------------------ reference API
    def set_variable(self, var, value=0, type='int', constant=False):
        """reference: regulargrid.py:85-95"""
        if isinstance(var, Symbol):
            var = var.name
        self.defined_variable[var] = Variable(var, value, type, constant)

    def calc_derivatives(self, max_order=1):
        """reference: regulargrid.py:97-104"""
        self.max_derivative_order = max_order
        self.set_order(self.order)
        for field in self.fields:
            field.populate_derivatives(max_order=max_order)

    def set_order(self, order):
        """reference: regulargrid.py:106-131"""
        for x in order:
            if x % 2 == 1:
                raise ValueError(str(x) + ' is not a valid order (require even integer)')
        self.order = order
        num_time_vars = max(self.order[0], self.max_derivative_order + 1)
        self.tp = Variable('tp', num_time_vars, 'int', True)
        self.time = [Variable(str(self.t) + str(k), k, 'int', False) for k in range(num_time_vars)]
        self.margin.value = self.order[1] // 2
        self.set_grid_size(self.grid_size)
        self.update_field_order()
        self._update_spacing()

    def get_time_step_limit(self):
        """reference: regulargrid.py:133-135"""
        h = min([sp.value for sp in self.spacing])
        return h / (3 ** 0.5 * self.c)

    def set_domain_size(self, size):
        self.size = size
        self._update_spacing()

    def set_grid_size(self, size):
        """dim = grid_size + 1 + 2*margin (reference: regulargrid.py:145-155)"""
        self.grid_size = size
        self.dim = [Variable('dim' + str(k + 1), size[k] + 1 + 2 * self.margin.value, 'int', True)
                    for k in range(self.dimension)]
        self._update_spacing()

    def update_field_order(self):
        if hasattr(self, 'fields'):
            for field in self.fields:
                field.set_order(self.order)

    def set_time_step(self, dt, tmax):
        """reference: regulargrid.py:165-172"""
        self.dt.value = dt
        self.ntsteps.value = int(tmax / dt)

    def set_switches(self, **kwargs):
        """reference: regulargrid.py:174-185"""
        for switch, value in kwargs.items():
            if switch not in self._switches:
                raise KeyError("Unsupported switch: ", switch)
            if not isinstance(value, bool):
                raise ValueError("Only boolean values allowed for switches")
            setattr(self, switch, value)
            if switch == 'double':
                self.real_t = 'double' if self.double else 'float'
                self._update_spacing()

    def _update_spacing(self):
        """reference: regulargrid.py:187-201"""
        self.spacing = [Variable('dx' + str(k + 1),
                                 self.size[k] / (self.dim[k].value - 1 - self.margin.value * 2),
                                 self.real_t, True) for k in range(self.dimension)]
        expr = self.order[0]
        for d in self.dim:
            expr *= d.value
        self.vec_size = Variable('vec_size', expr, 'int', True)

    def set_index(self, index):
        """reference: regulargrid.py:203-217"""
        if index is None:
            self.index = [Symbol('x' + str(k + 1)) for k in range(self.dimension)]
        else:
            self.index = index
        if hasattr(self, 'fields'):
            for field in self.fields:
                field.set_indices([self.t] + self.index)

    def set_params(self, **kwargs):
        """reference: regulargrid.py:219-224"""
        for param, value in kwargs.items():
            if param not in self._params:
                raise KeyError("Unsupported parameter: ", param)
            setattr(self, param, value)
            self.set_variable(param, value, self.real_t, True)

    def set_field_spacing(self):
        for field in self.fields:
            field.set_spacing(variable_to_symbol([self.dt] + self.spacing))

    def set_papi_events(self, events=[]):
        self._papi_events = events

    def get_all_variables(self):
        return self.dim + self.spacing + self.time + [self.tp, self.dt, self.margin, self.ntsteps] \
            + list(self.defined_variable.values())

    def create_const_dict(self):
        """reference: regulargrid.py:281-291"""
        self.const_dict = {Symbol(v.name): v.value for v in self.get_all_variables() if v.constant}

    # ------------------------------------------------------------------ PDE -> coefficients
    def _linear_coefficients(self, eq):
        """Read `lhs = sum_i coef_i * D_i` off one PDE: {DDerivative: coefficient expression}."""
        rhs = expand(eq.rhs)
        derivs = set(get_all_objects(rhs, DDerivative))
        coefs, rest = {}, rhs
        for d in derivs:
            c = rhs.coeff(d)
            coefs[d] = c
            rest = rest - c * d
        if expand(rest) != 0 or any(get_all_objects(c, DDerivative) for c in coefs.values()):
            raise NotImplementedError("only PDEs that are linear in the derivative symbols are supported")
        return coefs

    def _value(self, expr):
        """Numeric value of a coefficient expression with the defined constants substituted."""
        self.create_const_dict()
        v = expr.subs(self.const_dict) if hasattr(expr, 'subs') else expr
        return float(v)

    def solve_fd(self, equations):
        """reference: regulargrid.py:230-270.  The reference solves each FD-substituted equation
        for the newest time level with sympy; here the equation is only analysed: it must be
        `d2u/dt2 = sum_d w_d d2u/dx_d2` (second derivatives of the single field)."""
        if not len(self.fields) == len(equations):
            raise KeyError("Number of equations must be the same as number of fields. Number of fields is ",
                           len(self.fields))
        self.eq = list(equations)
        self.generic = False
        self.axis_weights = None
        try:
            self._solve_fd_acoustic()
        except NotImplementedError:
            # any other PDE system: derive the update like the reference does and compile it at run time
            self._solve_fd_generic()

    def _solve_fd_acoustic(self):
        """`d2u/dt2 = sum_d w_d d2u/dx_d2` of a single field: served by the fixed-function kernels."""
        if len(self.fields) != 1:
            raise NotImplementedError("fixed-function path: one field")
        field = self.fields[0]
        eq = self.eq[0]
        if not (isinstance(eq.lhs, DDerivative) and _same_field(eq.lhs.field, field) and eq.lhs.axis == 0
                and eq.lhs.order == 2):
            raise NotImplementedError("fixed-function path: d2u/dt2 = sum_d w_d d2u/dx_d2 only")
        weights = [0] * self.dimension
        for d, c in self._linear_coefficients(eq).items():
            if not _same_field(d.field, field) or d.order != 2 or d.axis == 0:
                raise NotImplementedError("fixed-function path: unsupported term %s" % d)
            weights[d.axis - 1] = weights[d.axis - 1] + c
        self.axis_weights = weights
        field.set_dt(eq.rhs)

    def _solve_fd_generic(self):
        """reference: regulargrid.py:230-270, verbatim in method: substitute every derivative symbol by its
        finite-difference expression of the grid's accuracy, solve each equation for the field at the newest time level
        with sympy, keep the symbolic kernel.  Lowered by `_build_params_generic` into CUDA source that the library
        compiles with NVRTC (include/opesci_b200.h: OPESCI_KIND_REGULAR_GENERIC)."""
        from sympy import solve
        if self.dimension != 3:
            raise NotImplementedError("B200 path: 3-D models only")
        if len(self.time) != 3:
            raise NotImplementedError("generic PDEs: three time levels (calc_derivatives(2), second order in time) -- "
                                      "the reference's RegularGrid maps the kernel onto _t0, _t1, _t2 (regulargrid.py:592-599)")
        t = self.t
        index_new = [t + 1 + (self.order[0] // 2 - 1)] + self.index
        simplify = max(self.order[1:]) <= 4
        for field, eq in zip(self.fields, self.eq):
            field.set_dt(eq.rhs)
            for deriv in get_all_objects(eq, DDerivative):
                eq = eq.subs(deriv, deriv.fd[deriv.max_accuracy])
            sols = solve(eq, field[index_new], simplify=simplify)
            if len(sols) != 1:
                raise NotImplementedError("equation for %s cannot be solved for the newest time level" % field.label)
            field.kernel = sols[0].subs({t: t - (self.order[0] // 2 - 1)})
        self.generic = True

    # ---- the reference's kernel transformations (regulargrid.py:329-342, 530-564, 601-604)
    def transform_kernel(self, field):
        kernel = field.kernel
        if self.expand:
            kernel = expand(kernel)
        if self.eval_const:
            self.create_const_dict()
            kernel = kernel.subs(self.const_dict)
        return kernel

    def kernel_sympy(self, field):
        tv = [Symbol(v.name) for v in self.time]
        return self.transform_kernel(field).xreplace({self.t + 1: tv[2], self.t: tv[1], self.t - 1: tv[0]})

    def second_initialisation_sympy(self, field):
        """regulargrid.py:547-556: the `-F[t-1]` term of the kernel is replaced by 2*v*dt and the sum halved."""
        v = symbols("v")
        kernel = self.transform_kernel(field)
        for arg in kernel.args:
            if str(arg).startswith("-") and str(self.t - 1) in str(arg):
                kernel = kernel.subs({arg: 0}, simultaneous=True)
                kernel = 0.5 * (kernel + 2 * v * self.dt)      # `self.dt` is a Variable, like in the reference: prints `v*dt`
        kernel = kernel.subs({self.t: Symbol(self.time[0].name)})
        for idx in self.index:
            kernel = kernel.subs(idx, Symbol('_' + idx.name))
        return kernel

    def generic_kernel_text(self):
        """The assignments the reference's generator would emit: (time-loop body, second initialisation body)."""
        tv = [Symbol(v.name) for v in self.time]
        loop = [Symbol('_' + x.name) for x in self.index]
        step = ['%s = %s;' % (ccode(f[[tv[2]] + self.index]), ccode(self.kernel_sympy(f))) for f in self.fields]
        init2 = ['%s = %s;' % (ccode(f[[tv[1]] + loop]), ccode(self.second_initialisation_sympy(f))) for f in self.fields]
        return step, init2

    def get_kernel_ai(self, fields=None):
        """(AI, AI_w, ADD, MUL, LOAD, STORE) of the expanded update as the reference counts it
        (regulargrid.py:293-327): one MUL per weighted neighbour, one ADD per extra term."""
        m = self.margin.value
        if getattr(self, 'generic', False):
            # one MUL per printed term with a literal, one ADD per extra term, all fields (regulargrid.py:293-327)
            self.create_const_dict()
            add = mul = 0
            loads = set()
            for f in self.fields:
                args = self.transform_kernel(f).args
                add += len(args) - 1
                mul += sum(1 for a in args if len(a.args) > 1)
                loads |= {str(i.base.label) for i in get_all_objects(self.transform_kernel(f), Indexed)}
            word = 8 if self.double else 4
            ai = float(add + mul) / (len(loads) + len(self.fields)) / word
            return (ai, ai * (add + mul) / max(add, mul, 1) / 2.0, add, mul, len(loads), len(self.fields))
        nterms = 1 + sum(2 * m for w in self.axis_weights if w != 0) + 1
        add, mul, load, store = nterms - 1, nterms - 1 + 1, 1, 1
        word = 8 if self.double else 4
        ai = float(add + mul) / (load + store) / word
        return (ai, ai * (add + mul) / max(add, mul) / 2.0, add, mul, load, store)

    # ------------------------------------------------------------------ lowering
    def _solution_variables(self, coords):
        """C variables visible to a printed solution expression (regulargrid.py:391-406, 681)."""
        rt = cexpr.DOUBLE if self.double else cexpr.FLOAT
        v = cexpr.Variables()
        ctypes_of = {'float': cexpr.FLOAT, 'double': cexpr.DOUBLE, 'int': cexpr.INT}
        for var in self.get_all_variables():
            if var.constant:
                v.scalar(var.name, ctypes_of[var.type], var.value)
        for d in range(self.dimension):
            n = self.dim[d].value
            v.axis('_' + self.index[d].name, cexpr.INT, d, np.arange(n))
            if coords is not None:
                v.axis(self.index[d].name, rt, d, coords[d])
        return v

    def _coordinates(self, shifts):
        """x = dx*(_x - m + shift) evaluated in real_t like the emitted
        `real_t x = dx1*(_x - 1.5F)` (staggeredgrid.py:637-641, regulargrid.py:680-681)."""
        rt = np.float64 if self.double else np.float32
        m = self.margin.value
        out = []
        for d in range(self.dimension):
            i = np.arange(self.dim[d].value)
            if shifts[d]:
                inner = i.astype(np.float32) - np.float32(m - 0.5)   # int - float literal -> float
            else:
                inner = i - m                                        # int - int -> int
            out.append(rt(self.spacing[d].value) * inner.astype(rt))
        return out

    def _common_params(self, kind, nfields, nlevels):
        p = abi.OpesciB200Params()
        p.struct_size = abi.ctypes.sizeof(abi.OpesciB200Params)
        p.kind, p.so, p.is_double = kind, self.order[1], int(self.double)
        for d in range(3):
            p.dim[d] = self.dim[d].value
            p.dx[d] = self.spacing[d].value
        p.ntsteps, p.nfields, p.nlevels = self.ntsteps.value, nfields, nlevels
        p.converge = int(bool(self.converge))
        p.dt = self.dt.value
        volume = 1.0
        for sp in self.spacing:
            volume *= sp.value
        p.volume_literal = float(literal(volume))
        return p

    def _field_spec(self, p, k, field, lo, hi, l2_lo, l2_hi, init_text, init_vars, final_text, final_vars, keep):
        fs = p.fields[k]
        for d in range(3):
            fs.lo[d], fs.hi[d] = lo[d], hi[d]
            fs.l2_lo[d], fs.l2_hi[d] = l2_lo[d], l2_hi[d]
        dims = [self.dim[d].value for d in range(3)]
        prog_i = cexpr.compile_expression(init_text, init_vars, dims)
        prog_f = cexpr.compile_expression(final_text, final_vars, dims)
        prog_i.fill(fs.init)
        prog_f.fill(fs.final_)
        keep += [prog_i, prog_f]

    def _residual_text(self, field, ti, tn, loop):
        """Printed `F[ti][_x][_y][_z] - sol(tn)` with the field access replaced by __F__."""
        placeholder = IndexedBase(str(field.label))[[ti] + loop]
        text = ccode(placeholder - field.sol.subs(self.t, tn))
        return text.replace(ccode(placeholder), '__F__')

    def _generic_source(self):
        """CUDA C++ for OPESCI_KIND_REGULAR_GENERIC: constants as the reference defines them at the top of
        opesci_execute (regulargrid.py:391-406), arrays as pointer-to-array casts (:474-496), one thread per point of
        the loops [m, dim-m)^3 (:566-590), the emitted assignments verbatim."""
        m = self.margin.value
        dims = [self.dim[d].value for d in range(3)]
        pitch = (dims[2] + 31) // 32 * 32
        rt = 'double' if self.double else 'float'
        step, init2 = self.generic_kernel_text()
        names = [str(f.label) for f in self.fields]
        consts = []
        for var in self.get_all_variables():
            if var.constant and var.name not in ('dim1', 'dim2', 'dim3'):
                consts.append('    const %s %s = %r;' % (var.type, var.name, var.value))
        casts = ['    real_t (*%s)[%d][%d][%d] = (real_t (*)[%d][%d][%d]) p%d;' % (n, dims[0], dims[1], pitch, dims[0], dims[1], pitch, k)
                 for k, n in enumerate(names)]
        params = ', '.join('real_t *p%d' % k for k in range(len(names)))

        def kernel(name, levels, x, y, z, body):
            return '\n'.join([
                'extern "C" __global__ void __launch_bounds__(256) %s(%s, %s)' % (name, params, ', '.join('int ' + l for l in levels)),
                '{',
                '    const int dim1 = %d, dim2 = %d, dim3 = %d;' % tuple(dims),
                '\n'.join(consts),
                '\n'.join(casts),
                '    const int %s = %d + (int)(blockIdx.x * blockDim.x + threadIdx.x);' % (z, m),
                '    const int %s = %d + (int)(blockIdx.y * blockDim.y + threadIdx.y);' % (y, m),
                '    const int %s = %d + (int)blockIdx.z;' % (x, m),
                '    if (%s >= dim3 - %d || %s >= dim2 - %d || %s >= dim1 - %d) return;' % (z, m, y, m, x, m),
                '\n'.join('    ' + line for line in body),
                '}', ''])
        idx = [i.name for i in self.index]
        tnames = [v.name for v in self.time]
        src = ['// generated by opesci_fd_b200.RegularGrid (generic PDE path); expressions as the reference prints them',
               'typedef %s real_t;' % rt,
               kernel('opesci_generic_step', tnames, idx[0], idx[1], idx[2], step),
               kernel('opesci_generic_init2', tnames[:2], '_' + idx[0], '_' + idx[1], '_' + idx[2], init2)]
        return '\n'.join(src)

    def _build_params_generic(self):
        keep = []
        m = self.margin.value
        nf = len(self.fields)
        if nf > abi.OPESCI_MAX_FIELDS:
            raise NotImplementedError("at most %d fields" % abi.OPESCI_MAX_FIELDS)
        p = self._common_params(abi.KIND_REGULAR_GENERIC, nf, len(self.time))
        p.free_surface = abi.FS_NONE
        source = self._generic_source().encode()
        p.generic_source = source
        keep.append(source)
        self.generic_source = source.decode()
        loop = [Symbol('_' + x.name) for x in self.index]
        coords = self._coordinates([False] * 3)
        ti = self.ntsteps.value % 2
        tn = self.dt.value * self.ntsteps.value
        dims = [self.dim[d].value for d in range(3)]
        for k, field in enumerate(self.fields):
            # level 0 := sol(t=0) over the WHOLE array with the integer loop indices as coordinates
            # (regulargrid.py:498-528); L2 on [m,dim-m) against sol(ntsteps*dt) (regulargrid.py:650-700)
            sol0 = field.sol.subs(self.t, 0)
            for idx in self.index:
                sol0 = sol0.subs(idx, Symbol('_' + idx.name))
            fvars = self._solution_variables(coords)
            fvars.field('__F__', cexpr.DOUBLE if self.double else cexpr.FLOAT)
            self._field_spec(p, k, field, [0] * 3, dims, [m] * 3, [n - m for n in dims],
                             ccode(sol0), self._solution_variables(None),
                             self._residual_text(field, ti, tn, loop), fvars, keep)
        return p, keep

    def build_params(self):
        if getattr(self, 'generic', False):
            return self._build_params_generic()
        if not getattr(self, 'axis_weights', None):
            raise RuntimeError("solve_fd() must be called before the model can be lowered")
        if self.order[0] != 2:
            raise NotImplementedError("time order %d" % self.order[0])
        keep = []
        m = self.margin.value
        p = self._common_params(abi.KIND_REGULAR_ACOUSTIC, 1, len(self.time))
        p.free_surface = abi.FS_NONE
        field = self.fields[0]
        # ---- update stencil u[t2] = -u[t0] + sum coef*u[t1][+-k] + centre*u[t1]
        # (regulargrid.py:592-619; literals as printed after expand + constant folding)
        a = central_weights(m, 2)
        dt2 = _frac(self.dt.value) ** 2
        centre = Fraction(2)
        for d in range(3):
            w = _frac(self._value(self.axis_weights[d])) if self.axis_weights[d] != 0 else Fraction(0)
            scale = dt2 * w / _frac(self.spacing[d].value) ** 2
            centre += scale * a[0]
            for k in range(1, m + 1):
                val = float(scale * a[k])
                p.ac_coef[d][k - 1] = literal(val) if w != 0 else 0.0
                # second initialisation: 0.5*(stencil + 2*v*dt) (regulargrid.py:550-556)
                p.ac_init_coef[d][k - 1] = literal(0.5 * val) if w != 0 else 0.0
        p.ac_centre = literal(float(centre))
        p.ac_init_centre = literal(0.5 * float(centre))
        rt = np.float64 if self.double else np.float32
        types = {'float': np.float32, 'double': np.float64}
        vvar = self.defined_variable.get('v')
        vval = types[vvar.type](vvar.value) if vvar is not None else rt(0)
        # `1.0F*v*dt`, left to right in C
        p.ac_init_const = float((np.float32(1.0) * vval) * types[self.dt.type](self.dt.value))
        # ---- analytic solution: level 0 := sol(t=0) over the WHOLE array with the integer loop
        # indices as coordinates (regulargrid.py:498-528); L2 on [m,dim-m) (regulargrid.py:650-700)
        loop = [Symbol('_' + x.name) for x in self.index]
        sol0 = field.sol.subs(self.t, 0)
        for idx in self.index:
            sol0 = sol0.subs(idx, Symbol('_' + idx.name))
        coords = self._coordinates([False] * 3)
        ti = self.ntsteps.value % 2
        tn = self.dt.value * self.ntsteps.value
        fvars = self._solution_variables(coords)
        fvars.field('__F__', cexpr.DOUBLE if self.double else cexpr.FLOAT)
        dims = [self.dim[d].value for d in range(3)]
        self._field_spec(p, 0, field, [0] * 3, dims, [m] * 3, [n - m for n in dims],
                         ccode(sol0), self._solution_variables(None),
                         self._residual_text(field, ti, tn, loop), fvars, keep)
        return p, keep
