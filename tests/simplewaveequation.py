"""Regular-grid wave-equation driver -- same entry point and CLI as the reference's
tests/simplewaveequation.py (reference: :9-81 model set-up, :84-133 `default`, :207-235 CLI),
running on the B200 library.  Usage:  python tests/simplewaveequation.py default -so 4 -x
"""
import os
import sys
from argparse import ArgumentParser, RawTextHelpFormatter
from os import path

sys.path.insert(0, path.dirname(path.dirname(path.abspath(__file__))))
from opesci_fd_b200 import *  # noqa: E402,F401,F403

_test_dir = path.join(path.dirname(path.abspath(__file__)), "src")


def simplewave3d(domain_size, grid_size, dt, tmax, output_vts=False, o_converge=True,
                 accuracy_order=[1, 2, 2, 2], omp=True, simd=False, ivdep=True, double=False, pluto=False,
                 filename='test.cpp', expand=True, eval_const=True, fission=False, verbose=True):
    """Scalar wave equation on a regular grid (reference: tests/simplewaveequation.py:9-81).

    NB the PDE sums d2/dx2 + d2/dy2 + d2/dy2 -- the y term twice, no z term -- exactly like
    the reference (simplewaveequation.py:76): parity means reproducing that."""
    if verbose:
        print('domain size: ' + str(domain_size))
        print('grid size: ' + str(grid_size))
        print('approximation order: ' + str(accuracy_order))
        print('dt: ' + str(dt))
        print('tmax: ' + str(tmax))

    MAIN_GRID = RegularField('MAIN_GRID', dimension=3)
    grid = RegularGrid(dimension=3, domain_size=domain_size, grid_size=grid_size, fields=[MAIN_GRID],
                       pluto=pluto, fission=fission)
    grid.set_time_step(dt, tmax)
    grid.set_switches(omp=omp, simd=simd, ivdep=ivdep, double=double, expand=expand,
                      eval_const=eval_const, output_vts=output_vts, converge=o_converge)
    t, x, y, z, const_c = symbols('_t x y z c')
    grid.set_index([x, y, z])
    grid.set_params(c=2, v=1)
    if verbose:
        print('require dt < ' + str(grid.get_time_step_limit()))
    mu = 10
    beta = 0.5
    Omega = pi * sqrt(2 * mu * beta)
    A = sqrt(2 * mu / beta)
    MAIN_GRID.set_analytic_solution(-A * sin(pi * x) * (sin(pi * y) - sin(pi * z)) * sin(Omega * t))
    grid.set_order(accuracy_order)
    grid.calc_derivatives(2)
    eq0 = Eq(MAIN_GRID.d[0][2], (const_c ** 2) * (MAIN_GRID.d[1][2] + MAIN_GRID.d[2][2] + MAIN_GRID.d[2][2]))
    grid.solve_fd([eq0])
    if verbose:
        print('Kernel AI')
        print('%.2f, %.2f (weighted), %d ADD, %d MUL, %d LOAD, %d STORE' % grid.get_kernel_ai())
    return grid


def default(compiler=None, execute=False, nthreads=1, accuracy_order=[2, 4, 4, 4], output=False,
            profiling=False, papi_events=[], pluto=False, tile=' ', fission=False, double=False,
            grid_size=(100, 100, 100), dt=0.002, tmax=1.0):
    """100^3 cells, 500 steps (reference: tests/simplewaveequation.py:84-133)."""
    domain_size = (1.0, 1.0, 1.0)
    os.makedirs(_test_dir, exist_ok=True)
    filename = path.join(_test_dir, 'regular3d.json')
    grid = simplewave3d(domain_size, grid_size, dt, tmax, accuracy_order=accuracy_order, o_converge=True,
                        omp=True, simd=False, ivdep=True, filename=filename, pluto=pluto, fission=fission,
                        double=double)
    grid.set_switches(output_vts=output, profiling=profiling)
    grid.set_papi_events(papi_events)
    out = None
    if compiler is None:
        grid.generate(filename)
    else:
        out = grid.compile(filename, compiler=compiler, shared=False)
    if execute:
        grid.execute(filename, compiler=compiler or 'g++', nthreads=nthreads)
        grid.convergence()
    return out


def main():
    p = ArgumentParser(description="Standalone testing script for the simple wave example",
                       formatter_class=RawTextHelpFormatter)
    p.add_argument('mode', choices=('default',), nargs='?', default='default')
    p.add_argument('-so', '--spatial_order', default=4, type=int, dest='so')
    p.add_argument('-c', '--compiler', default=None)
    p.add_argument('-x', '--execute', action='store_true', default=False)
    p.add_argument('-n', '--nthreads', type=int, default=1)
    p.add_argument('-o', '--output', action='store_true', default=False)
    p.add_argument('-p', '--profiling', action='store_true', default=False)
    p.add_argument('--papi-events', dest='papi_events', nargs='+', default=[])
    p.add_argument('--tile', default=None)
    p.add_argument('--pluto', action='store_true', default=False)
    p.add_argument('--fission', action='store_true', default=False)
    p.add_argument('--double', action='store_true', default=False)
    p.add_argument('--reference-l2', dest='reference_l2', action='store_true', default=False,
                   help='accumulate the L2 norms like the generated C++ does (serially, in real_t, in loop order): '
                        'prints the reference\'s own digits; slower (a serial chain)')
    args = p.parse_args()
    if args.reference_l2:
        os.environ['OPESCI_L2_REFERENCE'] = '1'
    print("Simple wave 3D example ")
    default(compiler=args.compiler, execute=args.execute, nthreads=args.nthreads,
            accuracy_order=[2, args.so, args.so, args.so], profiling=args.profiling, double=args.double)


if __name__ == "__main__":
    main()
