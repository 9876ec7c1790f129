"""CPU oracle for the nosh Newton-Krylov hot path.  TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  The product package
``nosh_b200`` never does (tests/test_boundary.py greps for it).
"""
from .oracle import OracleProblem, build_library, lib, num_threads, set_dot_parts, set_row_reverse  # noqa: F401
from . import meshgen  # noqa: F401
from . import continuation  # noqa: F401
from . import gmres  # noqa: F401
from . import fvm  # noqa: F401
