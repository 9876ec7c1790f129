"""nosh_b200 -- B200-native (sm_100a) Newton-Krylov hot path of nschloe/nosh.

The product is the C-ABI shared library ``libnosh_b200.so`` (include/nosh_b200.h) built from
``nosh_b200/csrc``; ``nosh_b200.api.Context`` is a thin ctypes front-end and
``nosh_b200/hostcpp`` the C++ mirror of the reference's classes.  No CPU fallback exists.
"""
from . import _lib  # noqa: F401
from .api import (Context, NoshError, morton_order, partition_range, read_mesh, renumber,  # noqa: F401
                  write_mesh)
from ._lib import (LAYOUT_CSR, LAYOUT_SELL32, MAT_DKEO, MAT_KEO, NO_TRANS, TRANS, CONJ_TRANS,  # noqa: F401
                   OP_JACOBIAN, OP_KEO, OP_KEOREG, PREC_NONE, PREC_KEOREG_AMG, SOLVER_MINRES, SOLVER_CG, SOLVER_GMRES, AMG_REUSE_NONE, AMG_REUSE_FULL, ARC_SCALING, ARC_HIT_BOUND,
                   FVM_VERTEX_NONE, FVM_VERTEX_EXP, FVM_VERTEX_EXP_LINEARIZED, FVM_DIRICHLET_NONE, FVM_DIRICHLET_IDENTITY,
                   FVM_DIRICHLET_ZERO, FVM_DIRICHLET_VALUE,
                   build)
