"""Named model constants.  Mirrors the reference interface opesci/variable.py:6-21
(`Variable(name, value=0, type='int', constant=False)`, a sympy Symbol carrying a value)."""
from sympy import Symbol

__all__ = ['Variable']


class Variable(Symbol):
    """A Symbol with a C type, a value and a const flag (reference: opesci/variable.py:6-21)."""

    def __new__(cls, name, *args, **kwargs):
        # uncached constructor: two grids in one process must not share Variable objects
        return Symbol.__xnew__(cls, str(name))

    def __init__(self, name, value=0, type='int', constant=False):
        self.type = type
        self.constant = constant
        self.value = value

    def __reduce_ex__(self, protocol):
        return (Variable, (self.name, self.value, self.type, self.constant))
