"""Compiler objects of the host front end.

The reference shells out to g++/clang++/icpc for every model (opesci/compilation.py:10-104).
On the B200 path nothing is compiled per model: every `Compiler.compile()` resolves to the
prebuilt CUDA library `csrc/libopesci_b200.so` (hand-written sm_100a kernels behind the C ABI
of include/opesci_b200.h).  The class names, constructor arguments and the
`compile(src, out=None, shared=True)` signature are kept so that drivers written for the
reference (`--compiler g++`) run unchanged.
"""
import os

from . import abi

__all__ = ['Compiler', 'GNUCompiler', 'ClangCompiler', 'IntelCompiler', 'B200Compiler']


def get_package_dir():
    return os.path.abspath(os.path.dirname(__file__))


class Compiler(object):
    """reference: opesci/compilation.py:10-48"""

    def __init__(self, cc, ld=None, cppargs=[], ldargs=[]):
        self._cc = os.environ.get('CC', cc)
        self._ld = os.environ.get('LDSHARED', ld)
        self._cppargs = cppargs
        self._ldargs = ldargs

    def compile(self, src, out=None, shared=True):
        """Return the library that executes the model described by `src`.

        A missing CUDA library is an error (no CPU fallback), like a failed compilation in
        the reference (opesci/compilation.py:37-46)."""
        lib = abi.CUDA_LIBRARY
        if not os.path.exists(lib):
            raise RuntimeError("Error during compilation:\nCUDA library %s is not built "
                               "(run __graft_entry__.build()).\nSource file: %s" % (lib, src))
        print("Compiled: %s (prebuilt sm_100a library; host compiler %s is not invoked)" % (lib, self._cc))
        return lib


class B200Compiler(Compiler):
    def __init__(self, cppargs=[], ldargs=[]):
        super(B200Compiler, self).__init__("nvcc", cppargs=cppargs, ldargs=ldargs)


class GNUCompiler(Compiler):
    """reference: opesci/compilation.py:51-66"""

    def __init__(self, cppargs=[], ldargs=[]):
        super(GNUCompiler, self).__init__("g++", cppargs=cppargs, ldargs=ldargs)


class ClangCompiler(Compiler):
    """reference: opesci/compilation.py:69-85"""

    def __init__(self, cppargs=[], ldargs=[]):
        super(ClangCompiler, self).__init__("clang++", cppargs=cppargs, ldargs=ldargs)


class IntelCompiler(Compiler):
    """reference: opesci/compilation.py:88-104"""

    def __init__(self, cppargs=[], ldargs=[]):
        super(IntelCompiler, self).__init__("icpc", cppargs=cppargs, ldargs=ldargs)
