"""Alias of solat_optics_b200.far_field_latrt (same public names as the reference module)."""
import sys as _sys

import solat_optics_b200.far_field_latrt as _impl

_sys.modules[__name__] = _impl
