"""C printer for solution expressions.

Mirrors the reference interface opesci/codeprinter.py (`ccode`, `ccode_eq`) and the three
printing rules that decide the numbers the reference computes with:
  * floats print in scientific notation with 15 significant digits and an `F` suffix, i.e.
    as FLOAT literals even in double mode (reference: opesci/codeprinter.py:46-63);
  * rationals print as `p.0F/q.0F` (codeprinter.py:24-30);
  * indexed accesses print as C arrays `U[t][x][y][z]` (codeprinter.py:14-22).
On the B200 path the printed text is never compiled: opesci_fd_b200/cexpr.py parses it and
evaluates it with C semantics, and `literal()` gives the value of a printed float literal.
"""
import numpy as np
from sympy import Eq
from sympy.printing.c import C89CodePrinter

__all__ = ['ccode', 'ccode_eq', 'literal', 'literal_text']


def literal_text(value):
    """The reference's decimal text of a float: 15 significant digits, trailing zeros of the
    mantissa stripped, scientific notation (reference: opesci/codeprinter.py:46-63)."""
    value = float(value)
    if value == 0.0:
        return '0.0'
    mant, exp = ('%.14e' % value).split('e')
    mant = mant.rstrip('0')
    if mant.endswith('.'):
        mant += '0'
    if int(exp) == 0:
        return mant              # mpmath's to_str omits a zero exponent: `1.988074375F`
    return '%se%d' % (mant, int(exp))


def literal(value):
    """float32 value of the `...F` literal the reference would print for `value`."""
    return np.float32(float(literal_text(value)))


class CodePrinter(C89CodePrinter):
    def _print_Indexed(self, expr):
        return self._print(expr.base.label) + ''.join('[%s]' % self._print(i) for i in expr.indices)

    def _print_Rational(self, expr):
        return '%d.0F/%d.0F' % (int(expr.p), int(expr.q))

    def _print_Float(self, expr):
        return literal_text(float(expr)) + 'F'

    def _print_Variable(self, expr):
        return expr.name

    def _print_DDerivative(self, expr):
        return expr.name


def ccode(expr, **settings):
    if isinstance(expr, Eq):
        return ccode_eq(expr)
    return CodePrinter(settings).doprint(expr, None)


def ccode_eq(eq, **settings):
    return CodePrinter(settings).doprint(eq.lhs, None) + ' = ' + CodePrinter(settings).doprint(eq.rhs, None)
