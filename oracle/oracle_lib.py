"""numpy-facing ctypes wrapper of the CPU oracle (oracle/solat_oracle.c).

TEST INFRASTRUCTURE ONLY (see the header of solat_oracle.c).  Importable only from
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libsolat_oracle.so")
_lib = None

_dp = ctypes.POINTER(ctypes.c_double)
_ip = ctypes.POINTER(ctypes.c_int)


def build(force=False):
    src = os.path.join(_HERE, "solat_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "CC=gcc"])
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
    return _lib


def _d(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(_dp)


class OracleBracketError(ValueError):
    """The reference raises ValueError from scipy.optimize.brentq in this situation."""


def _check(rc):
    if rc == -1:
        raise OracleBracketError("f(a) and f(b) must have different signs")
    if rc != 0:
        raise RuntimeError("oracle error %d" % rc)


def nf_rx_trace(p_rx, th_fwhp, n_scan, n_si=3.4, n_vac=1.0):
    out = np.zeros((20, n_scan * n_scan))
    p, pp = _d(p_rx)
    _check(lib().orc_nf_rx_trace(pp, ctypes.c_double(th_fwhp), ctypes.c_double(n_si), ctypes.c_double(n_vac),
                                 ctypes.c_int(n_scan), out.ctypes.data_as(_dp)))
    return out


def nf_source_trace(p_src, n_scan, n_si=3.4, n_vac=1.0):
    out = np.zeros((26, n_scan * n_scan))
    p, pp = _d(p_src)
    _check(lib().orc_nf_source_trace(pp, ctypes.c_double(n_si), ctypes.c_double(n_vac), ctypes.c_int(n_scan),
                                     out.ctypes.data_as(_dp)))
    return out


def nf_toy_trace(p_rx, y_ap, th_fwhp, n_scan, n_si=3.4, n_vac=1.0):
    out = np.zeros((32, n_scan * n_scan))
    p, pp = _d(p_rx)
    _check(lib().orc_nf_toy_trace(pp, ctypes.c_double(y_ap), ctypes.c_double(th_fwhp), ctypes.c_double(n_si),
                                  ctypes.c_double(n_vac), ctypes.c_int(n_scan), out.ctypes.data_as(_dp)))
    return out


def lyot_mask(n_scan, r_lyot=136.54 / 2):
    """scripts/holo_sim.py:168-171 on regrid's uniform index-space grid (rti:1413-1415,1430-1431)."""
    ax = np.linspace(-70, 70, n_scan)
    X2, Z2 = np.meshgrid(ax, ax)
    return (np.sqrt(np.ravel(X2) ** 2 + np.ravel(Z2) ** 2) <= r_lyot)


def nf_map(src_xz, y_ap, n_scan, rx_opl, rx_amp_masked, k_per_m, n_si=3.4, n_vac=1.0, nthreads=0):
    src, sp = _d(np.reshape(src_xz, (-1, 2)))
    npix = src.shape[0]
    ro, rop = _d(rx_opl)
    ra, rap = _d(rx_amp_masked)
    kk, kp = _d(np.atleast_1d(k_per_m))
    nf = kk.shape[0]
    amp = np.zeros((nf, npix))
    ph = np.zeros((nf, npix))
    _check(lib().orc_nf_map(sp, ctypes.c_int(npix), ctypes.c_double(y_ap), ctypes.c_double(n_si),
                            ctypes.c_double(n_vac), ctypes.c_int(n_scan), rop, rap, kp, ctypes.c_int(nf),
                            amp.ctypes.data_as(_dp), ph.ctypes.data_as(_dp), ctypes.c_int(nthreads)))
    return amp, ph


def ff_geo_vec(th_1=np.arctan(1 / 2), th2=(-np.pi / 2) - np.arctan(1 / 3), rotc_m=(0, -7.2, 0),
               tow_m=(0, -7.2, 100e3), rx_y_m=-1.32):
    return np.array([th_1, th2, rotc_m[0] * 1e3, rotc_m[1] * 1e3, rotc_m[2] * 1e3,
                     tow_m[0] * 1e3, tow_m[1] * 1e3, tow_m[2] * 1e3, rx_y_m * 1e3], dtype=np.float64)


def ff_ray_mirror_pts(geo, theta, phi, el, az):
    g, gp = _d(geo)
    th, thp = _d(theta)
    ph, php = _d(phi)
    n = th.shape[0]
    out = np.zeros((9, n * n))
    _check(lib().orc_ff_ray_mirror_pts(gp, thp, php, ctypes.c_int(n), ctypes.c_double(el), ctypes.c_double(az),
                                       out.ctypes.data_as(_dp)))
    return out


def ff_apert_fields(geo, theta, phi, focal_centre, el, az, pan_idx1=None, pan_idx2=None, pan1=None, pan2=None):
    g, gp = _d(geo)
    th, thp = _d(theta)
    ph, php = _d(phi)
    n = th.shape[0]
    if pan1 is None:
        pan1 = np.zeros((10, 1)); pan1[0, 0] = pan1[1, 0] = 1
        pan2 = pan1.copy()
        pan_idx1 = np.zeros(n * n, dtype=np.int32)
        pan_idx2 = np.zeros(n * n, dtype=np.int32)
    p1, p1p = _d(pan1)
    p2, p2p = _d(pan2)
    i1 = np.ascontiguousarray(pan_idx1, dtype=np.int32)
    i2 = np.ascontiguousarray(pan_idx2, dtype=np.int32)
    fc, fcp = _d(focal_centre)
    out = np.zeros((17, n * n))
    _check(lib().orc_ff_apert_fields(gp, thp, php, ctypes.c_int(n), fcp, i1.ctypes.data_as(_ip),
                                     i2.ctypes.data_as(_ip), p1p, ctypes.c_int(p1.shape[1]), p2p,
                                     ctypes.c_int(p2.shape[1]), ctypes.c_double(el), ctypes.c_double(az),
                                     out.ctypes.data_as(_dp)))
    return out


def ff_project(sim, data_x, data_y, k):
    s, sp = _d(sim)
    dx, dxp = _d(data_x)
    dy, dyp = _d(data_y)
    M = dx.shape[0]
    out = np.zeros((M, M), dtype=np.complex128)
    lib().orc_ff_project(sp, ctypes.c_int(s.shape[1]), dxp, dyp, ctypes.c_int(M), ctypes.c_double(k),
                         out.ctypes.data_as(_dp))
    return out


def ff_map(geo, azel, theta, phi, data_x, data_y, e_rec, k, two_pass=False, nthreads=0):
    g, gp = _d(geo)
    ae, aep = _d(np.reshape(azel, (-1, 2)))
    th, thp = _d(theta)
    ph, php = _d(phi)
    dx, dxp = _d(data_x)
    dy, dyp = _d(data_y)
    e = np.ascontiguousarray(e_rec, dtype=np.complex128)
    M = dx.shape[0]
    npix = ae.shape[0]
    amp = np.zeros(npix)
    pha = np.zeros(npix)
    _check(lib().orc_ff_map(gp, aep, ctypes.c_int(npix), thp, php, ctypes.c_int(th.shape[0]), dxp, dyp,
                            e.ctypes.data_as(_dp), ctypes.c_int(M), ctypes.c_double(k),
                            ctypes.c_int(1 if two_pass else 0), amp.ctypes.data_as(_dp),
                            pha.ctypes.data_as(_dp), ctypes.c_int(nthreads)))
    return amp, pha


def num_threads():
    return int(lib().orc_num_threads())
