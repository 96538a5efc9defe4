"""Eigenwave3D driver -- same entry point and CLI as the reference's tests/eigenwave3d.py
(reference: tests/eigenwave3d.py:9-146 model set-up, :149-198 `default`, :282-336 CLI), running
on the B200 library.  Usage:  python tests/eigenwave3d.py default -so 4 -c g++ -x

Differences to the reference driver: `--compiler` only selects a Compiler object (nothing is
compiled per model), `--pluto/--tile/--fission/--papi-events` are accepted and ignored (CPU loop
transformations); `read` mode follows the PATCHED reference semantics (SURVEY.md 0.8:
the unpatched reference produces NaN) and `--synthetic` replaces its absent data files.
"""
import os
import sys
from argparse import ArgumentParser, RawTextHelpFormatter
from os import path

sys.path.insert(0, path.dirname(path.dirname(path.abspath(__file__))))
from opesci_fd_b200 import *  # noqa: E402,F401,F403

_test_dir = path.join(path.dirname(path.abspath(__file__)), "src")


def eigenwave3d(domain_size, grid_size, dt, tmax, output_vts=False, o_converge=True,
                accuracy_order=[1, 2, 2, 2], omp=True, simd=False, ivdep=True, double=False, pluto=False,
                filename='test.cpp', read=False, expand=True, eval_const=True,
                rho_file='', vp_file='', vs_file='', fission=False, rho=1.0, vp=1.0, vs=0.5, verbose=True):
    """3-D elastic eigenwave in a box with six free surfaces (reference: tests/eigenwave3d.py:9-146)."""
    if verbose:
        print('domain size: ' + str(domain_size))
        print('grid size: ' + str(grid_size))
        print('approximation order: ' + str(accuracy_order))
        print('dt: ' + str(dt))
        print('tmax: ' + str(tmax))

    Txx = SField('Txx', dimension=3, direction=(1, 1))
    Tyy = SField('Tyy', dimension=3, direction=(2, 2))
    Tzz = SField('Tzz', dimension=3, direction=(3, 3))
    Txy = SField('Txy', dimension=3, direction=(1, 2))
    Tyz = SField('Tyz', dimension=3, direction=(2, 3))
    Txz = SField('Txz', dimension=3, direction=(1, 3))
    U = VField('U', dimension=3, direction=1)
    V = VField('V', dimension=3, direction=2)
    W = VField('W', dimension=3, direction=3)

    grid = StaggeredGrid(dimension=3, domain_size=domain_size, grid_size=grid_size,
                         stress_fields=[Txx, Tyy, Tzz, Txy, Tyz, Txz],
                         velocity_fields=[U, V, W], pluto=pluto, fission=fission)
    grid.set_time_step(dt, tmax)
    grid.set_switches(omp=omp, simd=simd, ivdep=ivdep, double=double, expand=expand,
                      eval_const=eval_const, output_vts=output_vts, converge=o_converge)

    rho_s, beta, lam, mu = symbols('rho beta lambda mu')
    t, x, y, z = symbols('_t x y z')
    grid.set_index([x, y, z])

    if read:
        grid.set_media_params(read=True, rho_file=rho_file, vp_file=vp_file, vs_file=vs_file)
    else:
        grid.set_media_params(read=False, rho=rho, vp=vp, vs=vs)
    if verbose:
        print('require dt < ' + str(grid.get_time_step_limit()))

    # eigenwave solution (reference: tests/eigenwave3d.py:88-108)
    Omega = pi * sqrt(2 * mu * beta)
    A = sqrt(2 * mu / beta)
    U.set_analytic_solution(cos(pi * x) * (sin(pi * y) - sin(pi * z)) * cos(Omega * t))
    V.set_analytic_solution(cos(pi * y) * (sin(pi * z) - sin(pi * x)) * cos(Omega * t))
    W.set_analytic_solution(cos(pi * z) * (sin(pi * x) - sin(pi * y)) * cos(Omega * t))
    Txx.set_analytic_solution(-A * sin(pi * x) * (sin(pi * y) - sin(pi * z)) * sin(Omega * t))
    Tyy.set_analytic_solution(-A * sin(pi * y) * (sin(pi * z) - sin(pi * x)) * sin(Omega * t))
    Tzz.set_analytic_solution(-A * sin(pi * z) * (sin(pi * x) - sin(pi * y)) * sin(Omega * t))
    Txy.set_analytic_solution(Float(0))
    Tyz.set_analytic_solution(Float(0))
    Txz.set_analytic_solution(Float(0))

    grid.set_order(accuracy_order)
    grid.calc_derivatives()

    # momentum equations
    eq1 = Eq(U.d[0][1], beta * (Txx.d[1][1] + Txy.d[2][1] + Txz.d[3][1]))
    eq2 = Eq(V.d[0][1], beta * (Txy.d[1][1] + Tyy.d[2][1] + Tyz.d[3][1]))
    eq3 = Eq(W.d[0][1], beta * (Txz.d[1][1] + Tyz.d[2][1] + Tzz.d[3][1]))
    # stress-strain equations
    eq4 = Eq(Txx.d[0][1], (lam + 2 * mu) * U.d[1][1] + lam * (V.d[2][1] + W.d[3][1]))
    eq5 = Eq(Tyy.d[0][1], (lam + 2 * mu) * V.d[2][1] + lam * (U.d[1][1] + W.d[3][1]))
    eq6 = Eq(Tzz.d[0][1], (lam + 2 * mu) * W.d[3][1] + lam * (U.d[1][1] + V.d[2][1]))
    eq7 = Eq(Txy.d[0][1], mu * (U.d[2][1] + V.d[1][1]))
    eq8 = Eq(Tyz.d[0][1], mu * (V.d[3][1] + W.d[2][1]))
    eq9 = Eq(Txz.d[0][1], mu * (U.d[3][1] + W.d[1][1]))
    grid.solve_fd([eq1, eq2, eq3, eq4, eq5, eq6, eq7, eq8, eq9])

    for dimension in (1, 2, 3):
        for side in (0, 1):
            grid.set_free_surface_boundary(dimension=dimension, side=side)

    if verbose:
        print('stress kernel AI')
        print('%.2f, %.2f (weighted), %d ADD, %d MUL, %d LOAD, %d STORE' % grid.get_stress_kernel_ai())
        print('velocity kernel AI')
        print('%.2f, %.2f (weighted), %d ADD, %d MUL, %d LOAD, %d STORE' % grid.get_velocity_kernel_ai())
        print('overall algorithm AI')
        print('%.2f, %.2f (weighted)' % grid.get_overall_kernel_ai())
    return grid


def default(compiler=None, execute=False, nthreads=1, accuracy_order=[2, 4, 4, 4], output=False,
            profiling=False, papi_events=[], pluto=False, tile=' ', fission=False, double=False,
            grid_size=(100, 100, 100), dt=0.002, tmax=1.0):
    """Eigenwave test case on a unit cube, 100^3 cells, 500 steps (reference: tests/eigenwave3d.py:149-198)."""
    domain_size = (1.0, 1.0, 1.0)
    os.makedirs(_test_dir, exist_ok=True)
    filename = path.join(_test_dir, 'eigenwave3d.json')
    grid = eigenwave3d(domain_size, grid_size, dt, tmax, accuracy_order=accuracy_order, o_converge=True,
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


def read_data(compiler=None, execute=False, nthreads=1, accuracy_order=[2, 4, 4, 4], output=False,
              profiling=False, papi_events=[], fission=False, synthetic=False):
    """Model initialisation from input files: heterogeneous rho / vp / vs, 200^3 arrays at so=4
    (reference: tests/eigenwave3d.py:201-228).  The reference's data files RHOhomogx200, VPhomogx200,
    VShomogx200 are not in its repository; `synthetic=True` (CLI: --synthetic) substitutes the random
    medium of SURVEY.md 8d config 5."""
    domain_size = (1.0, 1.0, 1.0)
    grid_size = (195, 195, 195)
    dt = 0.002
    tmax = 1.0
    os.makedirs(_test_dir, exist_ok=True)
    filename = path.join(_test_dir, 'eigenwave3d_read.json')
    grid = eigenwave3d(domain_size, grid_size, dt, tmax, read=True, accuracy_order=accuracy_order,
                       o_converge=False, omp=True, simd=False, ivdep=True, filename=filename,
                       rho_file='RHOhomogx200', vp_file='VPhomogx200', vs_file='VShomogx200', fission=fission)
    grid.set_switches(output_vts=output, profiling=profiling)
    grid.set_papi_events(papi_events)
    if synthetic:
        from opesci_fd_b200.util import synthetic_media
        grid.set_media_arrays(*synthetic_media([d.value for d in grid.dim]))
    if compiler is None:
        grid.generate(filename)
    else:
        grid.compile(filename, compiler=compiler, shared=False)
    if execute:
        grid.execute(filename, compiler=compiler or 'g++', nthreads=nthreads)
        grid.convergence()
    return grid


def converge_test(execute=True):
    """(2,4)-scheme convergence sweep h = 1/10 .. 1/80, dt ~ h^2 (reference: tests/eigenwave3d.py:247-279)."""
    domain_size = (1.0, 1.0, 1.0)
    s = 10
    c = 0.4 * s
    results = []
    tmp_dir = path.join(path.dirname(path.abspath(__file__)), 'tmp')    # the reference writes into ./tmp
    os.makedirs(tmp_dir, exist_ok=True)
    for _ in range(4):
        dt = c / (s ** 2)
        filename = path.join(tmp_dir, 'test3d_' + str(s) + '.json')
        grid = eigenwave3d(domain_size, (s, s, s), dt, 5.0, o_converge=True, accuracy_order=[2, 4, 4, 4],
                           filename=filename)
        if execute:
            grid.execute(filename)
            results.append((s, grid.convergence()))
            grid.free()
        s = s * 2
    return results


def main():
    ModeHelp = """Avalable testing modes:
default:   Eigenwave test case on a unit cube grid (100 x 100 x 100)

read:      Test for model intialisation from input file (RHOhomogx200, VPhomogx200, VShomogx200 in the
           working directory, or --synthetic)

converge:  Convergence test of the (2,4) scheme, which is 2nd order
           in time and 4th order in space. The test halves spacing
           starting from 0.1 and reduces dt by a factor of 4 for
           each step
"""
    p = ArgumentParser(description="Standalone testing script for the Eigenwave3D example",
                       formatter_class=RawTextHelpFormatter)
    p.add_argument('mode', choices=('default', 'read', 'converge', 'cx1'), nargs='?', default='default',
                   help=ModeHelp)
    p.add_argument('-so', '--spatial_order', default=4, type=int, dest='so',
                   help='order of the spatial discretisation to use, eg. 4 for 4th order in x,y,z')
    p.add_argument('-c', '--compiler', default=None, help='C++ Compiler name (kept for CLI compatibility)')
    p.add_argument('-x', '--execute', action='store_true', default=False,
                   help='Dynamically execute the model on the GPU')
    p.add_argument('-n', '--nthreads', type=int, default=1, help='Number of host threads (unused on GPU)')
    p.add_argument('-o', '--output', action='store_true', default=False, help='write U_<ti>.vts after every time step (output_vts switch)')
    p.add_argument('-p', '--profiling', action='store_true', default=False,
                   help='Print time-loop timing from CUDA events')
    p.add_argument('--papi-events', dest='papi_events', nargs='+', default=[], help='(accepted, ignored)')
    p.add_argument('--tile', default=None, help='(accepted, ignored)')
    p.add_argument('--pluto', action='store_true', default=False, help='(accepted, ignored)')
    p.add_argument('--fission', action='store_true', default=False, help='(accepted, ignored)')
    p.add_argument('--double', action='store_true', default=False, help='use double precision fields')
    p.add_argument('--synthetic', action='store_true', default=False,
                   help='read mode: use the synthetic random medium instead of the data files')
    p.add_argument('--reference-l2', dest='reference_l2', action='store_true', default=False,
                   help='accumulate the L2 norms like the generated C++ does (serially, in real_t, in loop order): '
                        'prints the reference\'s own digits; slower (a serial chain)')
    args = p.parse_args()
    if args.reference_l2:
        os.environ['OPESCI_L2_REFERENCE'] = '1'
    print("Eigenwave3D example (mode=%s)" % args.mode)

    if args.mode == 'default':
        default(compiler=args.compiler, execute=args.execute, nthreads=args.nthreads, output=args.output,
                accuracy_order=[2, args.so, args.so, args.so], profiling=args.profiling,
                double=args.double)
    elif args.mode == 'read':
        read_data(compiler=args.compiler, execute=args.execute, nthreads=args.nthreads, output=args.output,
                  accuracy_order=[2, args.so, args.so, args.so], profiling=args.profiling, synthetic=args.synthetic)
    elif args.mode == 'converge':
        converge_test()
    elif args.mode == 'cx1':
        raise NotImplementedError("cx1: Intel-cluster pragma comparison (SURVEY.md 2 item 18), CPU only")


if __name__ == "__main__":
    main()
