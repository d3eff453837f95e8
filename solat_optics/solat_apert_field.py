"""Alias of solat_optics_b200.solat_apert_field (same public names as the reference module)."""
import sys as _sys

import solat_optics_b200.solat_apert_field as _impl

_sys.modules[__name__] = _impl
