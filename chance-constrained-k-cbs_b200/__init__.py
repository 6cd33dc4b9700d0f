"""chance-constrained-k-cbs_b200 — B200-native belief propagation + chance-constraint checking
(the data-parallel hot path of CCK-CBS) behind a C ABI (include/cck_b200.h).

The directory name follows the repository spec and is not a valid Python identifier; import it
through the root-level ``cck_b200`` shim (``import cck_b200``).
"""
from .capi import *  # noqa: F401,F403
from . import capi  # noqa: F401
