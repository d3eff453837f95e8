"""Drop-in import name: `import solat_optics.ray_trace_int as rti`, `from solat_optics import ot_geo, ray_trace`
resolve to the B200 implementation in `solat_optics_b200` (put this repository ahead of the reference on
PYTHONPATH).  `ray_trace` is bound to the live `ray_trace_int` API, the one-word repair scripts/holo_sim.py:21
needs."""
from solat_optics_b200 import __version__  # noqa: F401
