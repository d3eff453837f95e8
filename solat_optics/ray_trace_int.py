"""Alias of solat_optics_b200.ray_trace_int (same public names as the reference module)."""
import sys as _sys

import solat_optics_b200.ray_trace_int as _impl

_sys.modules[__name__] = _impl
