"""Finite-difference weights and small expression helpers.

Mirrors the public names of the reference's opesci/util.py (`Deriv`, `Deriv_half`,
`get_all_objects`, `is_half`, `variable_to_symbol`, `IndexedBases`, `hf`), but the weights are
obtained differently: the reference inverts a symbolic Taylor matrix with sympy
(opesci/util.py:87-118, 136-236); here the same linear system is solved once in exact
rational arithmetic (`fd_weights`) and cached -- microseconds instead of seconds, and the
result is the same rational numbers (tests/test_frontend.py).
"""
from fractions import Fraction
from functools import lru_cache

from sympy import Basic, Indexed, IndexedBase, Rational, Symbol

hf = Rational(1, 2)

__all__ = ['hf', 'fd_weights', 'staggered_first_weights', 'central_weights', 'Deriv', 'Deriv_half',
           'get_all_objects', 'is_half', 'variable_to_symbol', 'IndexedBases', 'shift_grid',
           'shift_index']


@lru_cache(maxsize=None)
def fd_weights(nodes, order):
    """Exact weights w_j with  sum_j w_j f(x0 + nodes[j]*h) = h^order * f^(order)(x0) + O(h^p).

    `nodes` is a tuple of Fractions (offsets in units of h).  Solves the Vandermonde system
    sum_j w_j nodes[j]^p / p! = [p == order], p = 0..n-1, by Gauss-Jordan over the rationals.
    """
    n = len(nodes)
    fact = [Fraction(1)]
    for p in range(1, n):
        fact.append(fact[-1] * p)
    A = [[Fraction(x) ** p / fact[p] for x in nodes] + [Fraction(int(p == order))] for p in range(n)]
    for col in range(n):
        piv = next(r for r in range(col, n) if A[r][col] != 0)
        A[col], A[piv] = A[piv], A[col]
        inv = 1 / A[col][col]
        A[col] = [v * inv for v in A[col]]
        for r in range(n):
            if r != col and A[r][col] != 0:
                f = A[r][col]
                A[r] = [v - f * w for v, w in zip(A[r], A[col])]
    return tuple(A[r][n] for r in range(n))


def staggered_first_weights(m):
    """c_1..c_m of the staggered first derivative of order 2m:
    f'(x) ~ (1/h) sum_k c_k (f(x+(k-1/2)h) - f(x-(k-1/2)h))   (SURVEY.md 8a table)."""
    nodes = tuple(Fraction(2 * j - 1, 2) for j in range(-m + 1, m + 1))
    w = fd_weights(nodes, 1)
    return [w[m + k - 1] for k in range(1, m + 1)]


def central_weights(m, order):
    """a_0..a_m of the central derivative of accuracy 2m on integer nodes -m..m
    (symmetric for even `order`, antisymmetric for odd)."""
    nodes = tuple(Fraction(j) for j in range(-m, m + 1))
    w = fd_weights(nodes, order)
    return [w[m + k] for k in range(0, m + 1)]


def _rat(fr):
    return Rational(fr.numerator, fr.denominator)


def Deriv(U, index, k, d, n):
    """Central FD approximations of accuracy n along index k with spacing d
    (reference interface: opesci/util.py:136-173).  Returns [f, f', f'', ...] expressions."""
    m = n // 2
    result = []
    for order in range(0, n + 1):
        w = fd_weights(tuple(Fraction(j) for j in range(-m, m + 1)), order)
        expr = 0
        for j, wj in zip(range(-m, m + 1), w):
            if wj != 0:
                idx = list(index)
                idx[k] = idx[k] + j
                expr += _rat(wj) * U[tuple(idx)]
        result.append(expr / d ** order)
    return result


def Deriv_half(U, index, dimension, delta, order):
    """Staggered FD approximations on half-integer nodes, accuracy 2*order
    (reference interface: opesci/util.py:195-236).  Returns [f, f', ...] expressions."""
    n = 2 * order
    nodes = tuple(Fraction(j, 2) for j in range(-n + 1, n, 2))
    result = []
    for deriv in range(0, n):
        w = fd_weights(nodes, deriv)
        expr = 0
        for x, wj in zip(nodes, w):
            if wj != 0:
                idx = list(index)
                idx[dimension] = idx[dimension] + _rat(x)
                expr += _rat(wj) * U[tuple(idx)]
        result.append(expr / delta ** deriv)
    return result


def get_all_objects(expr, typ):
    """All sub-objects of `expr` that are instances of `typ` (reference: opesci/util.py:7-20)."""
    if isinstance(expr, typ):
        return [expr]
    if not isinstance(expr, Basic):
        return []
    found = []
    for arg in expr.args:
        found += get_all_objects(arg, typ)
    return found


def variable_to_symbol(variables):
    return [Symbol(v.name) for v in variables]


def IndexedBases(s):
    return tuple(IndexedBase(x) for x in s.split())


def is_half(expr):
    """True when the constant part of an index is integer + 1/2 (reference: opesci/util.py:239-247)."""
    zero = {x: 0 for x in expr.free_symbols}
    return not expr.subs(zero).is_Integer


def shift_grid(expr):
    """Drop the half from every staggered index (reference: opesci/util.py:250-267)."""
    if expr.is_Symbol or expr.is_Number:
        return expr
    if isinstance(expr, Indexed):
        return Indexed(expr.base, *[x - hf if is_half(x) else x for x in expr.indices])
    return expr.func(*[shift_grid(a) for a in expr.args])


def shift_index(expr, k, s):
    """Shift the k-th index of every field access by s (reference: opesci/util.py:270-292)."""
    if expr.is_Symbol or expr.is_Number:
        return expr
    if isinstance(expr, Indexed):
        idx = list(expr.indices)
        idx[k] += s
        return Indexed(expr.base, *idx)
    return expr.func(*[shift_index(a, k, s) for a in expr.args])


def synthetic_media(dims, seed=20261017):
    """Synthetic heterogeneous medium for `read` mode: rho, vp, vs i.i.d.
    uniform per cell from a counter-based generator (numpy Philox), rho in [1.0,1.5),
    vp in [1.0,1.5), vs in [0.4,0.7) -- keeps lambda = rho*(vp^2 - 2 vs^2) > 0 and mu > 0.
    Returns three float32 arrays of shape `dims` (the file layout of
    opesci_read_simple_binary_ptr: flat little-endian float32, dim1*dim2*dim3 values,
    reference: opesci/staggeredgrid.py:549-551)."""
    import numpy as np
    gen = np.random.Generator(np.random.Philox(seed))
    n = int(np.prod(dims))
    rho = (1.0 + 0.5 * gen.random(n, dtype=np.float32)).astype(np.float32).reshape(dims)
    vp = (1.0 + 0.5 * gen.random(n, dtype=np.float32)).astype(np.float32).reshape(dims)
    vs = (0.4 + 0.3 * gen.random(n, dtype=np.float32)).astype(np.float32).reshape(dims)
    return rho, vp, vs


def synthetic_media_planes(dims, plane0, nplanes, seed=20261017):
    """The same kind of medium as `synthetic_media`, generated plane by plane from a counter-based key
    (Philox keyed with (seed, global plane index)), so that every rank of a slab run can produce exactly
    the planes it stores -- halo planes included -- without materialising the global arrays.
    Returns rho, vp, vs of shape [nplanes][dim2][dim3] holding the global planes [plane0, plane0+nplanes)."""
    import numpy as np
    d2, d3 = int(dims[1]), int(dims[2])
    out = [np.empty((nplanes, d2, d3), dtype=np.float32) for _ in range(3)]
    for k in range(nplanes):
        gen = np.random.Generator(np.random.Philox(key=[seed, plane0 + k]))
        out[0][k] = (1.0 + 0.5 * gen.random((d2, d3), dtype=np.float32))
        out[1][k] = (1.0 + 0.5 * gen.random((d2, d3), dtype=np.float32))
        out[2][k] = (0.4 + 0.3 * gen.random((d2, d3), dtype=np.float32))
    return tuple(out)
