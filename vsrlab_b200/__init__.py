"""vsrlab_b200 — B200 (sm_100a) kernels and host scheduling for the Real-BasicVSR /
BasicVSR hot path of santurini/vsrlab.  The user-facing drop-in lives in the sibling
`vsrlab` package (same class paths as the reference); this package holds the C-ABI
library (`csrc/`, `include/vsrb200.h`), its ctypes binding and the forward scheduler.
"""
from ._lib import VsrbError, load  # noqa: F401
from .functional import (clear_caches, current_dtype, output_dtype, precision, set_output_dtype,  # noqa: F401
                         set_precision)

__all__ = ["VsrbError", "load", "set_precision", "precision", "current_dtype", "clear_caches", "set_output_dtype", "output_dtype"]
