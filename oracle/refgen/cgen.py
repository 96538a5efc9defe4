"""Minimal stand-in for the third-party `cgen` package (not installed here, no network).

TEST INFRASTRUCTURE ONLY.  The reference generator (opesci/regulargrid.py,
opesci/staggeredgrid.py, opesci/templates/*.py in /root/reference) builds its C++
through `cgen` AST nodes and finally calls ``str()`` on the top-level Module.  This
file implements just the node classes the reference touches, with the attribute
names its own helper relies on (opesci/cgen_wrapper.py:11-24 reads ``.contents``,
``.body`` and ``.text``).  It is written from the public cgen API description, it is
not a copy of cgen.
"""


def _lines(obj):
    """Flatten a node (or nested lists of nodes / strings) into source lines."""
    if obj is None:
        return []
    if isinstance(obj, (list, tuple)):
        out = []
        for o in obj:
            out += _lines(o)
        return out
    if isinstance(obj, Generable):
        return list(obj.generate())
    return [str(obj)]


class Generable(object):
    def generate(self):
        raise NotImplementedError

    def __str__(self):
        return "\n".join(self.generate())


class Line(Generable):
    def __init__(self, text=""):
        self.text = text

    def generate(self):
        yield str(self.text).rstrip("\n")


class Statement(Generable):
    def __init__(self, text):
        self.text = text

    def generate(self):
        yield str(self.text) + ";"


class Assign(Generable):
    def __init__(self, lvalue, rvalue):
        self.lvalue = lvalue
        self.rvalue = rvalue

    def generate(self):
        lhs = self.lvalue.inline() if isinstance(self.lvalue, Declarator) else str(self.lvalue)
        yield "%s = %s;" % (lhs, self.rvalue)


class Pragma(Generable):
    def __init__(self, value):
        self.value = value

    def generate(self):
        yield "#pragma %s" % self.value


class Define(Generable):
    def __init__(self, symbol, value):
        self.symbol = symbol
        self.value = value

    def generate(self):
        yield "#define %s %s" % (self.symbol, self.value)


class Include(Generable):
    def __init__(self, filename, system=True):
        self.filename = filename
        self.system = system

    def generate(self):
        if self.system:
            yield "#include <%s>" % self.filename
        else:
            yield '#include "%s"' % self.filename


# ---------------------------------------------------------------- declarators
class Declarator(Generable):
    def get_decl_pair(self):
        """-> (type string, declarator string)"""
        raise NotImplementedError

    def inline(self):
        t, d = self.get_decl_pair()
        return "%s %s" % (t, d)

    def generate(self):
        yield self.inline() + ";"


class Value(Declarator):
    def __init__(self, typename, name):
        self.typename = typename
        self.name = name

    def get_decl_pair(self):
        return str(self.typename), str(self.name)


class _Nested(Declarator):
    def __init__(self, subdecl):
        self.subdecl = subdecl


class Pointer(_Nested):
    def get_decl_pair(self):
        t, d = self.subdecl.get_decl_pair()
        return t, "*" + d


class Const(_Nested):
    def get_decl_pair(self):
        t, d = self.subdecl.get_decl_pair()
        return "const " + t, d


class ArrayOf(_Nested):
    def __init__(self, subdecl, count=None):
        _Nested.__init__(self, subdecl)
        self.count = count

    def get_decl_pair(self):
        t, d = self.subdecl.get_decl_pair()
        return t, "%s[%s]" % (d, "" if self.count is None else self.count)


class Initializer(Generable):
    def __init__(self, vdecl, data):
        self.vdecl = vdecl
        self.data = data

    def inline(self):
        return "%s = %s" % (self.vdecl.inline(), self.data)

    def generate(self):
        yield self.inline() + ";"


class InlineInitializer(Initializer):
    def generate(self):
        yield self.inline()


class FunctionDeclaration(_Nested):
    def __init__(self, subdecl, arg_decls):
        _Nested.__init__(self, subdecl)
        self.arg_decls = arg_decls

    def get_decl_pair(self):
        t, d = self.subdecl.get_decl_pair()
        return t, "%s(%s)" % (d, ", ".join(a.inline() for a in self.arg_decls))


class Extern(Generable):
    """extern "<language>" <declaration>"""

    def __init__(self, language, subdecl):
        self.language = language
        self.subdecl = subdecl

    def _prefix(self):
        return 'extern "%s" ' % self.language

    def inline(self):
        return self._prefix() + self.subdecl.inline()

    def generate(self):
        sub = list(self.subdecl.generate())
        sub[0] = self._prefix() + sub[0]
        return iter(sub)


class Struct(Declarator):
    def __init__(self, tpname, fields, declname=None):
        self.tpname = tpname
        self.fields = fields
        self.declname = declname

    def generate(self):
        yield "struct %s" % self.tpname
        yield "{"
        for ln in _lines(self.fields):
            yield "  " + ln
        yield "} %s;" % (self.declname or "")


# ---------------------------------------------------------------- containers
class Block(Generable):
    def __init__(self, contents=None):
        self.contents = list(contents) if contents is not None else []

    def generate(self):
        yield "{"
        for ln in _lines(self.contents):
            yield "  " + ln
        yield "}"

    def append(self, item):
        self.contents.append(item)

    def extend(self, items):
        self.contents.extend(items)


class Module(Block):
    def generate(self):
        for ln in _lines(self.contents):
            yield ln


class Loop(Generable):
    def __init__(self, body):
        self.body = body

    def intro_line(self):
        raise NotImplementedError

    def generate(self):
        yield self.intro_line()
        if isinstance(self.body, Block) and not isinstance(self.body, Module):
            for ln in self.body.generate():
                yield ln
        else:
            yield "{"
            for ln in _lines(self.body):
                yield "  " + ln
            yield "}"


def _inline_text(x):
    if isinstance(x, Initializer):
        return x.inline()
    if isinstance(x, Generable):
        return " ".join(x.generate())
    return str(x)


class For(Loop):
    def __init__(self, start, condition, update, body):
        Loop.__init__(self, body)
        self.start = start
        self.condition = condition
        self.update = update

    def intro_line(self):
        return "for (%s; %s; %s)" % (_inline_text(self.start), _inline_text(self.condition),
                                     _inline_text(self.update))


class IfDef(Generable):
    def __init__(self, condition, iflines, elselines):
        self.condition = condition
        self.iflines = iflines
        self.elselines = elselines

    def generate(self):
        yield "#ifdef %s" % self.condition
        for ln in _lines(self.iflines):
            yield ln
        yield "#else"
        for ln in _lines(self.elselines):
            yield ln
        yield "#endif"


class FunctionBody(Generable):
    def __init__(self, fdecl, body):
        self.fdecl = fdecl
        self.body = body

    def generate(self):
        yield self.fdecl.inline()
        for ln in self.body.generate():
            yield ln
