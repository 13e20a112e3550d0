"""oracle/ — TEST INFRASTRUCTURE ONLY.

CPU restatement of the mag2d hot path (``libmag2d_oracle.so``, plain C) and a ctypes window onto the
unmodified reference compiled into ``oracle/_ref`` (``libmag2d_ref_{parity,fast}.so``).

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this package — as the checker, never as the thing measured or shipped.  The product
(``mag2d_b200``) must not import it and fails loudly when its CUDA library is missing.
"""
from .pyoracle import Oracle, Oracle3, Orc3Grid, OrcGrid, build_oracle  # noqa: F401
from .pyref import Ref3D, RefHarness, ref_available, REF_DIR  # noqa: F401
