#!/usr/bin/env python
"""Make the (Python-2-only) reference generator importable under Python 3.12 / sympy 1.14.

TEST INFRASTRUCTURE ONLY -- used to build `oracle/_ref/` and `tests/golden/` in the
development container.  Nothing here is on the product path and nothing here runs on
the GPU box (/root/reference does not exist there).

The reference sources are NOT copied into this repository.  This script copies
`/root/reference/opesci` and the two driver scripts into a scratch directory
(default /tmp/opesci_py3) and applies the mechanical edits listed in SURVEY.md 8c
in place there:

  1. print statements -> print(); file() -> open(); drop `from __builtin__ import str`;
     implicit relative imports -> absolute `opesci.` imports
  2. integer `/` -> `//` where the result is used as an int
  3. sympy.printing.ccode.CCodePrinter -> sympy.printing.c.C89CodePrinter; print our
     Variable symbols by name (clash with sympy.codegen.ast.Variable)
  4. get_all_objects: recurse through Basic, not Expr (Eq is not an Expr any more)
  5. Eq(expr) -> Eq(expr, 0)
  6. Field/Media: per-(class, name) instance cache so that sympy's `func(*args)`
     rebuilds keep the python-side attributes (.staggered, .bc, ...)
  7. '\\partial ' names -> plain names; map() -> list(map()); dict.values() -> list
  8. `cgen` -> the stub next to this file

No numerical statement of the reference is touched.
"""
import os
import re
import shutil
import sys

REF = os.environ.get("OPESCI_REFERENCE", "/root/reference")
HERE = os.path.dirname(os.path.abspath(__file__))


def _sub(text, pattern, repl, count=0, flags=0, must=True, what=""):
    new, n = re.subn(pattern, repl, text, count=count, flags=flags)
    if must and n == 0:
        raise RuntimeError("shim pattern did not match: %s (%s)" % (pattern, what))
    return new


def _fix_prints(text):
    out = []
    lines = text.split("\n")
    i = 0
    while i < len(lines):
        ln = lines[i]
        m = re.match(r"^(\s*)print (.*)$", ln)
        if m and not ln.lstrip().startswith("#"):
            body = m.group(2)
            if body.count('"""') == 1:  # multi-line string literal
                while True:
                    i += 1
                    body += "\n" + lines[i]
                    if '"""' in lines[i]:
                        break
            out.append("%sprint(%s)" % (m.group(1), body))
        else:
            out.append(ln)
        i += 1
    return "\n".join(out)


def shim(dst):
    if os.path.exists(dst):
        shutil.rmtree(dst)
    os.makedirs(dst)
    shutil.copytree(os.path.join(REF, "opesci"), os.path.join(dst, "opesci"))
    os.makedirs(os.path.join(dst, "drivers"))
    for f in ("eigenwave3d.py", "simplewaveequation.py"):
        shutil.copy(os.path.join(REF, "tests", f), os.path.join(dst, "drivers", f))
    shutil.copy(os.path.join(HERE, "cgen.py"), os.path.join(dst, "cgen.py"))

    def edit(rel, fn):
        p = os.path.join(dst, rel)
        with open(p) as fh:
            t = fh.read()
        t2 = fn(t)
        with open(p, "w") as fh:
            fh.write(t2)

    # ---- package __init__: absolute imports, no versioneer
    def f_init(t):
        t = re.sub(r"^from (\w+) import \*", r"from opesci.\1 import *", t, flags=re.M)
        t = t.replace("from ._version import get_versions", "")
        t = t.replace("__version__ = get_versions()['version']", "__version__ = 'py3-shim'")
        t = t.replace("del get_versions", "")
        return t
    edit("opesci/__init__.py", f_init)

    def absolutise(t):
        mods = ["grid", "variable", "codeprinter", "derivative", "util", "fields",
                "compilation", "regulargrid", "staggeredgrid"]
        for m in mods:
            t = re.sub(r"^from %s import" % m, "from opesci.%s import" % m, t, flags=re.M)
        t = re.sub(r"^import cgen_wrapper as cgen", "import opesci.cgen_wrapper as cgen", t, flags=re.M)
        t = re.sub(r"^from templates import", "from opesci.templates import", t, flags=re.M)
        t = re.sub(r"^from __builtin__ import str\n", "", t, flags=re.M)
        t = re.sub(r"^import includes$", "from opesci.templates import includes", t, flags=re.M)
        t = re.sub(r"^from regular3d_tmpl import", "from opesci.templates.regular3d_tmpl import", t, flags=re.M)
        t = t.replace("with file(", "with open(")
        return _fix_prints(t)

    for rel in ["opesci/grid.py", "opesci/compilation.py", "opesci/regulargrid.py",
                "opesci/staggeredgrid.py", "opesci/fields.py", "opesci/util.py",
                "opesci/codeprinter.py", "opesci/cgen_wrapper.py", "opesci/derivative.py",
                "opesci/variable.py", "opesci/templates/regular3d_tmpl.py",
                "opesci/templates/staggered3d_tmpl.py", "opesci/templates/includes.py",
                "drivers/eigenwave3d.py", "drivers/simplewaveequation.py"]:
        edit(rel, absolutise)

    # ---- util.py: integer division, Basic recursion
    def f_util(t):
        t = _sub(t, r"from sympy import Expr, ", "from sympy import Basic, Expr, ")
        t = _sub(t, r"if not isinstance\(expr, Expr\):", "if not isinstance(expr, Basic):")
        t = t.replace("range(-n/2, n/2+1)", "range(-(n//2), n//2+1)")
        return t
    edit("opesci/util.py", f_util)

    # ---- fields.py
    def f_fields(t):
        t = _sub(t, r"Deriv_half\(self, l, k, d, n/2\)", "Deriv_half(self, l, k, d, n//2)")
        t = t.replace("range(self.order[d]/2-1)", "range(self.order[d]//2-1)")
        t = _sub(t, r"eq = Eq\(expr\)\n", "eq = Eq(expr, 0)\n")
        t = _sub(t, r"eq = Eq\(self\[idx\]\)\n", "eq = Eq(self[idx], 0)\n")
        t = _sub(t, r"eq1 = Eq\(self\[idx\]\)\n", "eq1 = Eq(self[idx], 0)\n")
        t = _sub(t, r"name = ''\.join\(\['\\partial ', self\.label\.name, '/\\partial ', str\(index\)\]\)",
                 "name = 'D_' + self.label.name + '_' + str(index) + '_' + str(order)")
        # instance cache: sympy >= 1.? rebuilds IndexedBase through func(*args)
        cache = (
            "    _shim_cache = {}\n\n"
            "    def __new__(typ, name, *args, **kwargs):\n"
            "        key = (typ.__name__, str(name))\n"
            "        if not kwargs and key in typ._shim_cache:\n"
            "            return typ._shim_cache[key]\n"
            "        obj = IndexedBase.__new__(typ, name)\n"
            "        if kwargs:\n"
            "            typ._shim_cache[key] = obj\n"
            "        return obj\n")
        t = _sub(t, r"    def __new__\(typ, name, \*\*kwargs\):\n        obj = IndexedBase.__new__\(typ, name\)\n        return obj\n",
                 lambda m: cache)
        # __init__ is re-run on cached instances by python when __new__ returns one:
        # the reference already guards set() with len(kwargs); super().__init__() is harmless
        return t
    edit("opesci/fields.py", f_fields)

    # ---- codeprinter.py
    def f_cp(t):
        t = _sub(t, r"from sympy\.printing\.ccode import CCodePrinter",
                 "from sympy.printing.c import C89CodePrinter as CCodePrinter")
        t = _sub(t, r"args = map\(ccode, expr\.args\)", "args = list(map(ccode, expr.args))")
        t = _sub(t, r"(    def _print_Indexed\(self, expr\):)",
                 "    def _print_Variable(self, expr):\n        return expr.name\n\n\\1")
        return t
    edit("opesci/codeprinter.py", f_cp)

    # ---- regulargrid.py
    def f_rg(t):
        t = _sub(t, r"self\.margin\.value = self\.order\[1\]/2", "self.margin.value = self.order[1]//2")
        t = t.replace("(self.order[0]/2-1)", "(self.order[0]//2-1)")
        t = _sub(t, r"\+ self\.defined_variable\.values\(\)", "+ list(self.defined_variable.values())")
        return t
    edit("opesci/regulargrid.py", f_rg)

    # ---- staggeredgrid.py
    def f_sg(t):
        t = t.replace("(self.order[0]/2-1)", "(self.order[0]//2-1)")
        t = _sub(t, r"self\.dimension\*\(self\.dimension-1\)/2", "self.dimension*(self.dimension-1)//2")
        return t
    edit("opesci/staggeredgrid.py", f_sg)

    # ---- variable.py / derivative.py: Symbol subclasses carrying python attributes.
    # (`name` is a slot-less attribute on modern Symbol: assigning it is rejected)
    def f_dd(t):
        t = t.replace("        self.name = name\n", "")
        return t
    edit("opesci/derivative.py", f_dd)
    return dst


if __name__ == "__main__":
    d = sys.argv[1] if len(sys.argv) > 1 else "/tmp/opesci_py3"
    print(shim(d))
