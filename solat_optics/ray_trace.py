"""`ray_trace` bound to the live tracer API (reference scripts/holo_sim.py:21 imports this name but calls the
ray_trace_int interface)."""
import sys as _sys

import solat_optics_b200.ray_trace_int as _impl

_sys.modules[__name__] = _impl
