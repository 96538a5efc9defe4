"""opesci_fd_b200 -- B200-native execution of opesci-fd's time-stepping hot path.

The host front end keeps the reference's API (reference: opesci/__init__.py:1-8 re-exports
the same names): `StaggeredGrid`, `RegularGrid`, `SField`, `VField`, `RegularField`, `Media`,
`Variable`, `DDerivative`, `Deriv`, `Deriv_half`, `ccode` plus the sympy names its drivers use.
`grid.execute()` / `grid.convergence()` run on hand-written sm_100a CUDA kernels through the
C ABI of include/opesci_b200.h (opesci_fd_b200/csrc).  There is no CPU fallback.
"""
from .variable import *  # noqa: F401,F403
from .fields import *  # noqa: F401,F403
from .derivative import *  # noqa: F401,F403
from .staggeredgrid import *  # noqa: F401,F403
from .codeprinter import *  # noqa: F401,F403
from .util import *  # noqa: F401,F403
from .regulargrid import *  # noqa: F401,F403
from sympy import symbols, Eq, sqrt, pi, cos, sin, Float  # noqa: F401

__version__ = '0.1.0'
