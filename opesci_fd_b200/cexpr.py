"""Typed evaluation of the C expressions the reference prints for analytic solutions.

The reference pastes `ccode(solution)` into its init and L2 loops
(reference: opesci/staggeredgrid.py:647-654, 931-935; opesci/regulargrid.py:519-522, 685-689)
and lets the C++ compiler + libm evaluate it per cell.  To reproduce those values bit for bit
without generating code, this module

  1. parses the printed C expression (C precedence, left associativity),
  2. types every node the way C++ does (int / float / double; `F` literals are float; the
     <cmath> functions visible in the global namespace take and return double),
  3. evaluates every maximal sub-tree that depends on at most ONE loop coordinate on the
     host -- numpy IEEE arithmetic in the node's own type, `math.sin/cos/...` (the same libm
     the reference binary links) for function calls -- into a constant or a 1-D table, and
  4. emits the remaining multi-coordinate `+ - * /` tree as a postfix program that the
     device executes in IEEE double (include/opesci_b200.h: OpesciSolProgram).

Heterogeneous (`read`) mode: the printed solution contains per-cell media accesses such as
`sqrt(beta[_x][_y][_z]*mu[_x][_y][_z])` (reference: opesci/staggeredgrid.py:648-653).  They become
OP_MEDIA operands; `float`-typed per-cell arithmetic is the double operation followed by OP_ROUNDF
(double rounding is innocuous for + - * / sqrt from 24 to 53 bits), and sqrt / cos / sin of a per-cell
argument run on the device.
"""
import math
import re

import numpy as np

from . import abi

INT, FLOAT, DOUBLE = 0, 1, 2
_NP = {INT: np.int64, FLOAT: np.float32, DOUBLE: np.float64}

_CONSTANTS = {
    "M_PI": 3.14159265358979323846, "M_SQRT2": 1.41421356237309504880,
    "M_SQRT1_2": 0.70710678118654752440, "M_E": 2.7182818284590452354,
    "M_PI_2": 1.57079632679489661923, "M_PI_4": 0.78539816339744830962,
    "M_1_PI": 0.31830988618379067154, "M_2_PI": 0.63661977236758134308,
    "M_LN2": 0.69314718055994530942, "M_LN10": 2.30258509299404568402,
    "M_LOG2E": 1.4426950408889634074, "M_LOG10E": 0.43429448190325182765,
    "M_2_SQRTPI": 1.12837916709551257390,
}
_FUNCS = {"sin": math.sin, "cos": math.cos, "tan": math.tan, "sqrt": math.sqrt, "exp": math.exp,
          "log": math.log, "fabs": math.fabs, "sinh": math.sinh, "cosh": math.cosh,
          "tanh": math.tanh, "asin": math.asin, "acos": math.acos, "atan": math.atan,
          "pow": math.pow}

_TOKEN = re.compile(r"\s*(?:(\d+\.\d*(?:[eE][+-]?\d+)?|\.\d+(?:[eE][+-]?\d+)?|\d+[eE][+-]?\d+|\d+)([fFlL]?)"
                    r"|([A-Za-z_][A-Za-z_0-9]*)|(.))")


class Node(object):
    __slots__ = ("kind", "args", "value", "name", "ctype", "deps", "data")

    def __init__(self, kind, args=(), value=None, name=None):
        self.kind, self.args, self.value, self.name = kind, list(args), value, name
        self.ctype = None   # INT / FLOAT / DOUBLE
        self.deps = None    # frozenset of axes (0,1,2); 'F' marks the field operand
        self.data = None    # numpy scalar or 1-D array for deps <= 1 axis


def tokenize(text):
    pos, out = 0, []
    while pos < len(text):
        m = _TOKEN.match(text, pos)
        if not m:
            break
        pos = m.end()
        if m.group(1) is not None:
            out.append(("num", m.group(1), m.group(2)))
        elif m.group(3) is not None:
            out.append(("id", m.group(3), None))
        elif m.group(4) is not None and m.group(4).strip():
            out.append(("op", m.group(4), None))
    return out


class Parser(object):
    def __init__(self, text):
        self.toks = tokenize(text)
        self.i = 0

    def peek(self):
        return self.toks[self.i] if self.i < len(self.toks) else ("end", None, None)

    def take(self, val=None):
        t = self.peek()
        if val is not None and t[1] != val:
            raise SyntaxError("expected %r, got %r" % (val, t[1]))
        self.i += 1
        return t

    def parse(self):
        n = self.additive()
        if self.peek()[0] != "end":
            raise SyntaxError("trailing tokens: %r" % (self.peek(),))
        return n

    def additive(self):
        n = self.multiplicative()
        while self.peek()[0] == "op" and self.peek()[1] in "+-":
            op = self.take()[1]
            n = Node("bin", [n, self.multiplicative()], name=op)
        return n

    def multiplicative(self):
        n = self.unary()
        while self.peek()[0] == "op" and self.peek()[1] in "*/":
            op = self.take()[1]
            n = Node("bin", [n, self.unary()], name=op)
        return n

    def unary(self):
        t = self.peek()
        if t[0] == "op" and t[1] == "-":
            self.take()
            return Node("neg", [self.unary()])
        if t[0] == "op" and t[1] == "+":
            self.take()
            return self.unary()
        return self.primary()

    def primary(self):
        t = self.take()
        if t[0] == "num":
            text, suffix = t[1], t[2]
            if suffix in ("f", "F"):
                return Node("lit", value=np.float32(float(text)), name=FLOAT)
            if re.match(r"^\d+$", text):
                return Node("lit", value=np.int64(int(text)), name=INT)
            return Node("lit", value=np.float64(float(text)), name=DOUBLE)
        if t[0] == "id":
            if self.peek()[1] == "(":
                self.take("(")
                args = []
                if self.peek()[1] != ")":
                    args.append(self.additive())
                    while self.peek()[1] == ",":
                        self.take(",")
                        args.append(self.additive())
                self.take(")")
                return Node("call", args, name=t[1])
            return Node("var", name=t[1])
        if t[0] == "op" and t[1] == "(":
            n = self.additive()
            self.take(")")
            return n
        raise SyntaxError("unexpected token %r" % (t,))


class Variables(object):
    """Name -> (ctype, deps, data).  deps is a frozenset of axes; data a scalar or 1-D array."""

    def __init__(self):
        self.table = {}

    def scalar(self, name, ctype, value):
        self.table[name] = (ctype, frozenset(), _NP[ctype](value))

    def axis(self, name, ctype, axis, values):
        self.table[name] = (ctype, frozenset([axis]), np.asarray(values, dtype=_NP[ctype]))

    def field(self, name, ctype):
        self.table[name] = (ctype, frozenset(["F"]), None)

    def media(self, name, ctype, media_id):
        """per-cell array operand `name` (text `name[_x][_y][_z]` is rewritten to `name__cell`)"""
        self.table[name + "__cell"] = (ctype, frozenset(["C%d" % media_id]), None)


def _per_cell(deps):
    return any(isinstance(d, str) for d in deps)


_DEVICE_CALLS = {"sqrt": abi.OP_SQRT, "cos": abi.OP_COS, "sin": abi.OP_SIN}


def _cast(data, ctype):
    return None if data is None else np.asarray(data).astype(_NP[ctype])[()]


def _annotate(n, variables):
    """Bottom-up: C type, coordinate dependencies and (when <= 1 axis) the host value."""
    for a in n.args:
        _annotate(a, variables)
    if n.kind == "lit":
        n.ctype, n.deps, n.data = n.name, frozenset(), n.value
    elif n.kind == "var":
        if n.name in _CONSTANTS:
            n.ctype, n.deps, n.data = DOUBLE, frozenset(), np.float64(_CONSTANTS[n.name])
        elif n.name in variables.table:
            n.ctype, n.deps, n.data = variables.table[n.name]
        else:
            raise NameError("unknown identifier %r in solution expression" % n.name)
    elif n.kind == "neg":
        a = n.args[0]
        n.ctype, n.deps = a.ctype, a.deps
        n.data = None if a.data is None else -a.data
    elif n.kind == "bin":
        a, b = n.args
        n.ctype = max(a.ctype, b.ctype)   # usual arithmetic conversions: int < float < double
        n.deps = a.deps | b.deps
        if len(n.deps) <= 1 and not _per_cell(n.deps):
            x, y = _cast(a.data, n.ctype), _cast(b.data, n.ctype)
            with np.errstate(all="ignore"):
                if n.name == "+":
                    n.data = x + y
                elif n.name == "-":
                    n.data = x - y
                elif n.name == "*":
                    n.data = x * y
                else:
                    if n.ctype == INT:
                        raise NotImplementedError("integer division in a solution expression")
                    n.data = x / y
            n.data = np.asarray(n.data, dtype=_NP[n.ctype])[()]
    elif n.kind == "call":
        if n.name not in _FUNCS:
            raise NotImplementedError("function %r in a solution expression" % n.name)
        n.ctype = DOUBLE   # ::sin(double) etc.: <cmath> puts only the double versions in ::
        n.deps = frozenset().union(*[a.deps for a in n.args])
        if _per_cell(n.deps) and "F" not in n.deps and n.name in _DEVICE_CALLS and len(n.args) == 1:
            return   # evaluated per cell on the device
        if len(n.deps) > 1 or "F" in n.deps:
            raise NotImplementedError("%s() of more than one coordinate: not separable" % n.name)
        fn = _FUNCS[n.name]
        args = [np.asarray(_cast(a.data, DOUBLE), dtype=np.float64) for a in n.args]
        shape = np.broadcast(*args).shape
        if shape == ():
            n.data = np.float64(fn(*[float(a) for a in args]))
        else:
            bargs = [np.broadcast_to(a, shape) for a in args]
            n.data = np.array([fn(*[float(b[i]) for b in bargs]) for i in range(shape[0])],
                              dtype=np.float64)


class ProgramBuilder(object):
    def __init__(self, dims):
        self.dims = dims
        self.instr = []
        self.tables = []   # (axis, float64 array)

    def emit(self, n):
        if "F" in n.deps and n.kind == "var":
            self.instr.append((abi.OP_FIELD, 0, 0.0))
            return
        if n.kind == "var" and _per_cell(n.deps):
            self.instr.append((abi.OP_MEDIA, int(next(iter(n.deps))[1:]), 0.0))
            return
        if len(n.deps) == 0:
            self.instr.append((abi.OP_CONST, 0, float(np.float64(n.data))))
            return
        if len(n.deps) == 1 and not _per_cell(n.deps):
            axis = next(iter(n.deps))
            tab = np.ascontiguousarray(np.broadcast_to(np.asarray(n.data, dtype=np.float64),
                                                       (self.dims[axis],)), dtype=np.float64)
            for k, (ax, t) in enumerate(self.tables):
                if ax == axis and np.array_equal(t.view(np.int64), tab.view(np.int64)):
                    self.instr.append((abi.OP_TABLE, k, 0.0))
                    return
            self.tables.append((axis, tab))
            self.instr.append((abi.OP_TABLE, len(self.tables) - 1, 0.0))
            return
        # multi-coordinate node: must be double arithmetic to run on the device VM
        if n.kind == "neg":
            self.emit(n.args[0])
            self.instr.append((abi.OP_NEG, 0, 0.0))
        elif n.kind == "bin":
            if n.ctype == INT:
                raise NotImplementedError("multi-coordinate integer %s" % n.name)
            self.emit(n.args[0])
            self.emit(n.args[1])
            self.instr.append(({"+": abi.OP_ADD, "-": abi.OP_SUB, "*": abi.OP_MUL,
                                "/": abi.OP_DIV}[n.name], 0, 0.0))
            if n.ctype == FLOAT:
                self.instr.append((abi.OP_ROUNDF, 0, 0.0))
        elif n.kind == "call" and n.name in _DEVICE_CALLS:
            self.emit(n.args[0])
            self.instr.append((_DEVICE_CALLS[n.name], 0, 0.0))
        else:
            raise NotImplementedError("cannot lower node %r" % n.kind)


class Program(object):
    """Host-side program; keeps the table arrays alive while ctypes points at them."""

    def __init__(self, instr, tables):
        self.instr, self.tables = instr, tables
        if len(instr) > abi.OPESCI_MAX_PROG or len(tables) > abi.OPESCI_MAX_TABLES:
            raise NotImplementedError("solution expression too large for the device program")

    def fill(self, cprog):
        cprog.n_instr = len(self.instr)
        cprog.n_tables = len(self.tables)
        for k, (axis, tab) in enumerate(self.tables):
            cprog.table_axis[k] = axis
            cprog.table[k] = tab.ctypes.data_as(abi.POINTER(abi.c_double))
        for k, (op, arg, val) in enumerate(self.instr):
            cprog.instr[k].op, cprog.instr[k].arg, cprog.instr[k].value = op, arg, val

    def evaluate(self, x, y, z, fieldval=0.0, media=None):
        """Pure-python execution of the program at one cell (tests only)."""
        st, idx = [], (x, y, z)
        for op, arg, val in self.instr:
            if op == abi.OP_MEDIA:
                st.append(np.float64(media[arg][x, y, z]))
                continue
            if op in (abi.OP_SQRT, abi.OP_COS, abi.OP_SIN):
                st[-1] = np.float64({abi.OP_SQRT: math.sqrt, abi.OP_COS: math.cos, abi.OP_SIN: math.sin}[op](float(st[-1])))
                continue
            if op == abi.OP_ROUNDF:
                st[-1] = np.float64(np.float32(st[-1]))
                continue
            if op == abi.OP_TABLE:
                st.append(np.float64(self.tables[arg][1][idx[self.tables[arg][0]]]))
            elif op == abi.OP_CONST:
                st.append(np.float64(val))
            elif op == abi.OP_FIELD:
                st.append(np.float64(fieldval))
            elif op == abi.OP_NEG:
                st[-1] = -st[-1]
            else:
                b = st.pop()
                a = st.pop()
                st.append({abi.OP_ADD: a + b, abi.OP_SUB: a - b, abi.OP_MUL: a * b,
                           abi.OP_DIV: a / b if op == abi.OP_DIV else None}[op])
        return st[-1] if st else np.float64(0.0)


def compile_expression(text, variables, dims):
    """C expression text -> Program (see module docstring)."""
    text = re.sub(r"\b([A-Za-z_][A-Za-z_0-9]*)\[_x\]\[_y\]\[_z\]", r"\1__cell", text)
    tree = Parser(text).parse()
    _annotate(tree, variables)
    pb = ProgramBuilder(dims)
    pb.emit(tree)
    return Program(pb.instr, pb.tables)
