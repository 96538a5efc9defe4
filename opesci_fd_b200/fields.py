"""Field objects of the host front end.

Mirrors the reference interface of opesci/fields.py (`SField`, `VField`, `RegularField`,
`Media`): same constructor keywords, same attributes (`dimension`, `direction`, `staggered`,
`d[axis][order]`, `sol`, `bc`) and the same staggering rules (fields.py:175-190, 270-292).
What is NOT here is the symbolic kernel / boundary-condition derivation (fields.py:98-138,
192-261, 294-381): on the B200 path those kernels are fixed-function CUDA, so a field only
has to describe itself; the grid reads the PDE coefficients off the equations instead.
"""
from sympy import IndexedBase

from .derivative import DDerivative
from .util import Deriv, Deriv_half

__all__ = ['SField', 'VField', 'Media', 'RegularField']


class Field(IndexedBase):
    """Base class (reference: opesci/fields.py:8-167)."""

    _registry = {}

    def __new__(typ, name, *args, **kwargs):
        # sympy rebuilds IndexedBase objects through func(*args); hand back the registered
        # instance then, so python-side attributes survive symbolic manipulation
        key = (typ.__name__, str(name))
        if not kwargs and key in Field._registry:
            return Field._registry[key]
        obj = IndexedBase.__new__(typ, name)
        if kwargs:
            Field._registry[key] = obj
        return obj

    def __init__(self, *args, **kwargs):
        if kwargs:
            self.set(**kwargs)

    def set(self, dimension, staggered):
        self.dimension = dimension
        self.staggered = staggered
        self.bc = [[None] * 2 for _ in range(dimension + 1)]
        self.sol = None

    def set_analytic_solution(self, function):
        """Exact solution used for initialisation and the L2 test (fields.py:38-44)."""
        self.sol = function

    def set_indices(self, indices):
        self.indices = indices

    def set_spacing(self, spacing):
        self.spacing = spacing

    def set_order(self, order):
        self.order = order

    def calc_derivative(self, l, k, d, n, order_of_derivative):
        return Deriv_half(self, l, k, d, n // 2)[order_of_derivative]

    def populate_derivatives(self, max_order=1):
        """d[axis][order] = DDerivative with .fd[accuracy] expressions (fields.py:71-96)."""
        self.d = [[None] * (max_order + 1) for _ in range(self.dimension + 1)]
        for d in range(self.dimension + 1):
            index = self.indices[d]
            for order in range(1, max_order + 1):
                name = 'D%d_%s_%s' % (order, self.label.name, str(index))
                deriv = DDerivative(name, index, order, self.order[d], field=self, axis=d)
                for accuracy in range(2, self.order[d] + 2, 2):
                    deriv.fd[accuracy] = self.calc_derivative(self.indices, d, self.spacing[d],
                                                              accuracy, order)
                self.d[d][order] = deriv

    def set_dt(self, dt):
        self.dt = dt


class VField(Field):
    """Velocity component: staggered in time and along its own direction (fields.py:169-190)."""

    def set(self, dimension, direction):
        self.direction = direction
        staggered = [False] * (dimension + 1)
        staggered[0] = True
        staggered[direction] = True
        Field.set(self, dimension, staggered)


class SField(Field):
    """Stress component: normal stresses unstaggered, shear stresses staggered along both of
    their directions, never in time (fields.py:264-292)."""

    def set(self, dimension, direction):
        self.direction = direction
        staggered = [False] * (dimension + 1)
        if direction[0] != direction[1]:
            for d in direction:
                staggered[d] = True
        Field.set(self, dimension, staggered)


class RegularField(Field):
    """Unstaggered scalar field with central differences (fields.py:409-419)."""

    def __init__(self, *args, **kwargs):
        if kwargs:
            Field.set(self, kwargs['dimension'], [0] * (kwargs['dimension'] + 1))

    def set(self, dimension, staggered=None):
        Field.set(self, dimension, [0] * (dimension + 1))

    def calc_derivative(self, l, k, d, n, order_of_derivative):
        return Deriv(self, l, k, d, n)[order_of_derivative]


class Media(IndexedBase):
    """Spatially varying medium parameter (fields.py:384-406)."""

    def __new__(typ, name, **kwargs):
        return IndexedBase.__new__(typ, name)

    def __init__(self, *args, **kwargs):
        if kwargs:
            self.set(**kwargs)

    def set(self, dimension, staggered, index):
        self.dimension = dimension
        self.staggered = staggered
        self.index = index
