"""Derivative placeholders used to write PDEs.  Mirrors opesci/derivative.py:6-26."""
from sympy import Symbol

__all__ = ['DDerivative']


class DDerivative(Symbol):
    """Symbol standing for d^order F / d var^order (reference: opesci/derivative.py:6-26).

    `fd[accuracy]` holds the finite-difference expression of that accuracy.  In addition to
    the reference attributes, `field` and `axis` record what is differentiated so that
    `solve_fd` can read the PDE coefficients straight off the equations.
    """

    def __new__(cls, name, *args, **kwargs):
        return Symbol.__xnew__(cls, str(name))

    def __init__(self, name, var, order, max_accuracy, field=None, axis=None):
        self.var = var
        self.order = order
        self.max_accuracy = max_accuracy
        self.fd = [None] * (max_accuracy + 1)
        self.field = field
        self.axis = axis
