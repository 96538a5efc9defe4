"""PDE systems outside the fixed-function kernels, written against the front-end API only (RegularField, RegularGrid,
`d[axis][order]`, `solve_fd`): the SAME function builds the model with the reference's `opesci` package (oracle/refgen/
make_ref.py -> golden fixtures from the reference's own generated C++) and with `opesci_fd_b200` (tests/test_generic.py ->
NVRTC-compiled sm_100a kernels).  Reference for the API: tests/simplewaveequation.py:9-81, opesci/regulargrid.py:230-270.
"""
from sympy import Eq, pi, sin, cos, sqrt, symbols


def build(pkg, name, domain_size, grid_size, dt, tmax, accuracy_order, double=False, o_converge=True, **switches):
    """pkg: module exposing RegularField / RegularGrid (the reference's `opesci` or `opesci_fd_b200`)."""
    t, x, y, z, c = symbols('_t x y z c')
    if name == 'damped':
        # d2u/dt2 = c^2 (uxx + uyy + uzz) - a du/dt + b du/dx - k u      (one field; first derivatives in t and x, a reaction term)
        u = pkg.RegularField('U', dimension=3)
        fields = [u]
    elif name == 'coupled':
        # d2u/dt2 = c^2 (uxx + uyy + uzz) + g w ;  d2w/dt2 = q^2 (wxx + wzz) + b dw/dy - g u      (two coupled fields)
        u = pkg.RegularField('U', dimension=3)
        w = pkg.RegularField('W', dimension=3)
        fields = [u, w]
    else:
        raise KeyError(name)
    grid = pkg.RegularGrid(dimension=3, domain_size=domain_size, grid_size=grid_size, fields=fields)
    grid.set_time_step(dt, tmax)
    grid.set_switches(omp=True, simd=False, ivdep=True, double=double, expand=True, eval_const=True,
                      output_vts=False, converge=o_converge, **switches)
    grid.set_index([x, y, z])
    grid.set_params(c=1.5, v=0.75)
    a, b, k, g, q = symbols('a b k g q')
    grid.set_variable('a', 0.3, 'float', True)
    grid.set_variable('b', 0.2, 'float', True)
    grid.set_variable('k', 1.25, 'float', True)
    grid.set_variable('g', 0.5, 'float', True)
    grid.set_variable('q', 0.8, 'float', True)
    # (not solutions of the PDEs: they define the initial state -- evaluated, like the reference does, with the integer
    #  loop indices as coordinates, regulargrid.py:498-528 -- and the function the L2 output is measured against)
    u.set_analytic_solution(cos(pi * x) * (cos(pi * y) - cos(2 * pi * z) / 2) * cos(3 * t) + sin(pi * x) * sin(2 * t))
    if name == 'coupled':
        w.set_analytic_solution(cos(2 * pi * x) * cos(pi * y) * cos(pi * z) * cos(2 * t) / 4)
    grid.set_order(accuracy_order)
    grid.calc_derivatives(2)
    if name == 'damped':
        eqs = [Eq(u.d[0][2], (c ** 2) * (u.d[1][2] + u.d[2][2] + u.d[3][2]) - a * u.d[0][1] + b * u.d[1][1] - k * u[t, x, y, z])]
    else:
        eqs = [Eq(u.d[0][2], (c ** 2) * (u.d[1][2] + u.d[2][2] + u.d[3][2]) + g * w[t, x, y, z]),
               Eq(w.d[0][2], (q ** 2) * (w.d[1][2] + w.d[3][2]) + b * w.d[2][1] - g * u[t, x, y, z])]
    grid.solve_fd(eqs)
    grid.get_kernel_ai()      # as the reference driver does (tests/simplewaveequation.py:78-80); it also builds grid.const_dict
    return grid
