"""ctypes binding of libsolat_b200.so (C ABI: include/solat_b200.h).

There is no CPU fallback: if the library is missing or no sm_100 device is visible the
calls raise.  (The library itself can be *loaded* without a GPU, which is what the CPU-only
tests check: every symbol of include/solat_b200.h must be exported.)
"""
import ctypes
import os
import threading

import numpy as np

from . import _build

c_dp = ctypes.POINTER(ctypes.c_double)
c_ip = ctypes.POINTER(ctypes.c_int)
c_sp = ctypes.POINTER(ctypes.c_short)
c_vp = ctypes.c_void_p
c_d = ctypes.c_double
c_i = ctypes.c_int

ST_BRACKET, ST_NOCONV, ST_NAN, ST_RANGE = 1, 2, 4, 8


class SolatError(RuntimeError):
    pass


class NfGeom(ctypes.Structure):
    _fields_ = [("n_si", c_d), ("n_vac", c_d), ("th_fwhp", c_d), ("y_ap", c_d)]


class FfGeom(ctypes.Structure):
    _fields_ = [("th_1", c_d), ("th2", c_d), ("rotc_mm", c_d * 3), ("src_mm", c_d * 3), ("foc_y_mm", c_d)]


class PanelLayoutC(ctypes.Structure):
    _fields_ = [("rect", c_dp), ("model_col", c_ip), ("n", c_i)]


# name -> (restype, argtypes); mirrors include/solat_b200.h one to one
SIGNATURES = {
    "solat_abi_version": (c_i, []),
    "solat_last_error": (ctypes.c_char_p, []),
    "solat_device_count": (c_i, []),
    "solat_ctx_create": (c_i, [c_i, ctypes.POINTER(c_vp)]),
    "solat_ctx_destroy": (None, [c_vp]),
    "solat_ctx_info": (c_i, [c_vp, c_ip, c_ip, c_ip, ctypes.POINTER(ctypes.c_longlong)]),
    "solat_launch_count": (ctypes.c_longlong, []),
    "solat_fp64_peak": (c_i, [c_vp, c_i, c_dp]),
    "solat_selftest_math": (c_i, [c_vp, c_dp, c_i, c_dp]),
    "solat_nf_rx_trace": (c_i, [c_vp, ctypes.POINTER(NfGeom), c_dp, c_dp, c_i, c_dp, c_i, c_dp, c_ip]),
    "solat_nf_source_trace": (c_i, [c_vp, ctypes.POINTER(NfGeom), c_dp, c_dp, c_i, c_dp, c_i, c_dp, c_ip]),
    "solat_nf_toy_trace": (c_i, [c_vp, ctypes.POINTER(NfGeom), c_dp, c_dp, c_i, c_dp, c_i, c_dp, c_ip]),
    "solat_nf_rx_fields": (c_i, [c_vp, ctypes.POINTER(NfGeom), c_dp, c_i, c_dp, c_i, c_dp, c_i, c_dp, c_dp, c_ip]),
    "solat_regrid_bicubic": (c_i, [c_vp, c_dp, c_i, c_i, c_dp]),
    "solat_nf_plan_create": (c_i, [c_vp, ctypes.POINTER(NfGeom), c_dp, c_i, c_dp, c_i, c_dp, c_dp, c_dp, c_i,
                                   ctypes.POINTER(c_vp)]),
    "solat_nf_plan_create_multi": (c_i, [c_vp, ctypes.POINTER(NfGeom), c_dp, c_i, c_dp, c_i, c_dp, c_dp, c_i, c_dp, c_i,
                                         ctypes.POINTER(c_vp)]),
    "solat_nf_plan_destroy": (None, [c_vp]),
    "solat_nf_map_host": (c_i, [c_vp, c_dp, c_i, c_dp, c_dp, c_ip]),
    "solat_nf_map_device": (c_i, [c_vp, c_vp, c_i, c_vp, c_vp, c_vp]),
    "solat_nf_map_device_st": (c_i, [c_vp, c_vp, c_i, c_vp, c_vp, c_vp, c_vp]),
    "solat_nf_plan_active_rays": (c_i, [c_vp]),
    "solat_nf_plan_channels": (c_i, [c_vp]),
    "solat_nf_plan_status": (c_i, [c_vp, c_ip]),
    "solat_ff_ray_mirror_pts": (c_i, [c_vp, ctypes.POINTER(FfGeom), c_dp, c_i, c_dp, c_i, c_d, c_d, c_dp, c_ip]),
    "solat_ff_apert_fields": (c_i, [c_vp, ctypes.POINTER(FfGeom), c_dp, c_i, c_dp, c_i, c_dp, c_ip, c_ip, c_dp, c_i,
                                    c_dp, c_i, c_d, c_d, c_dp, c_ip]),
    "solat_ff_project": (c_i, [c_vp, c_dp, c_i, c_dp, c_dp, c_i, c_d, c_dp]),
    "solat_ff_plan_create": (c_i, [c_vp, ctypes.POINTER(FfGeom), c_dp, c_i, c_dp, c_i, c_dp, c_dp, c_dp, c_i, c_d,
                                   ctypes.POINTER(c_vp)]),
    "solat_ff_plan_destroy": (None, [c_vp]),
    "solat_ff_map_host": (c_i, [c_vp, c_dp, c_i, c_dp, c_dp, c_ip]),
    "solat_ff_map_device": (c_i, [c_vp, c_vp, c_i, c_vp, c_vp, c_vp]),
    "solat_ff_map_device_st": (c_i, [c_vp, c_vp, c_i, c_vp, c_vp, c_vp, c_vp]),
    "solat_ff_plan_status": (c_i, [c_vp, c_ip]),
    "solat_ff_map_mirror_xy_host": (c_i, [c_vp, c_dp, c_i, c_dp, c_ip]),
    "solat_ff_map_panels_host": (c_i, [c_vp, c_dp, c_i, c_sp, c_sp, c_dp, c_i, c_dp, c_i, c_dp, c_dp, c_ip]),
    "solat_ff_map_panel_layout_host": (c_i, [c_vp, c_dp, c_i, ctypes.POINTER(PanelLayoutC), ctypes.POINTER(PanelLayoutC), c_dp, c_i,
                                             c_dp, c_i, c_dp, c_dp, c_ip]),
}

_lib = None
_lock = threading.Lock()
_contexts = {}


def lib_path():
    return _build.LIB


def load():
    """Load the shared library (building it first if the sources are newer and nvcc exists)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = lib_path()
        if _build.stale():
            try:
                _build.build()
            except Exception as exc:  # no nvcc on this box and no prebuilt library
                if not os.path.exists(path):
                    raise SolatError(
                        "libsolat_b200.so is missing and could not be built (%s). The CUDA library is "
                        "mandatory: this package has no CPU fallback." % exc
                    )
        lib = ctypes.CDLL(path)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError here = ABI mismatch, fail loudly
            fn.restype = res
            fn.argtypes = args
        if lib.solat_abi_version() != 1:
            raise SolatError("libsolat_b200.so ABI version mismatch")
        _lib = lib
        return lib


def check(rc):
    if rc != 0:
        msg = load().solat_last_error()
        raise SolatError("libsolat_b200 error %d: %s" % (rc, msg.decode() if msg else "?"))


def dptr(a):
    return a.ctypes.data_as(c_dp)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """One per device (solat_ctx)."""

    def __init__(self, device=0):
        lib = load()
        h = c_vp()
        check(lib.solat_ctx_create(int(device), ctypes.byref(h)))
        self.handle = h
        self.device = int(device)
        sm, ma, mi, hbm = c_i(), c_i(), c_i(), ctypes.c_longlong()
        check(lib.solat_ctx_info(h, ctypes.byref(sm), ctypes.byref(ma), ctypes.byref(mi), ctypes.byref(hbm)))
        self.sm_count, self.cc, self.hbm_bytes = sm.value, (ma.value, mi.value), hbm.value

    def fp64_peak_tflops(self, repeats=5):
        v = c_d()
        check(load().solat_fp64_peak(self.handle, int(repeats), ctypes.byref(v)))
        return v.value

    def close(self):
        if self.handle:
            load().solat_ctx_destroy(self.handle)
            self.handle = None


def context(device=None):
    """Process-wide context cache.  device=None -> LOCAL_RANK (torchrun) or 0."""
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", "0"))
    with _lock:
        ctx = _contexts.get(device)
    if ctx is None:
        ctx = Context(device)
        with _lock:
            _contexts[device] = ctx
    return ctx


def selftest_math(x, device=None):
    """(10, n) accuracy probe of the device fast-math primitives; see include/solat_b200.h."""
    x = as_f64(x)
    out = np.empty((10, x.size))
    check(load().solat_selftest_math(context(device).handle, dptr(x), x.size, dptr(out)))
    return out


def launch_count():
    return int(load().solat_launch_count())


def nf_geom(tele_geo):
    return NfGeom(float(tele_geo.n_si), float(tele_geo.n_vac), float(tele_geo.th_fwhp), float(tele_geo.y_ap))


def ff_geom(tele_geo):
    g = FfGeom()
    g.th_1 = float(tele_geo.th_1)
    g.th2 = float(tele_geo.th2)
    g.rotc_mm[:] = [float(tele_geo.x_rotc) * 1e3, float(tele_geo.y_rotc) * 1e3, float(tele_geo.z_rotc) * 1e3]
    g.src_mm[:] = [float(tele_geo.x_tow) * 1e3, float(tele_geo.y_tow) * 1e3, float(tele_geo.z_tow) * 1e3]
    g.foc_y_mm = float(tele_geo.rx_y) * 1e3
    return g


class RayStatusError(ValueError):
    """Raised in strict mode where the reference's scipy.optimize.brentq raises ValueError
    ("f(a) and f(b) must have different signs"): a ray's root lies outside the reference bracket."""


def raise_on_status(st, strict):
    if strict and (int(st) & ST_BRACKET):
        raise RayStatusError("f(a) and f(b) must have different signs")
