"""ctypes wrapper of tests/host_shadow/shadow.cpp (the product's device math compiled for the
host).  Test infrastructure: lets the CPU-only suite check the exact per-ray source the kernels
run against the golden fixtures."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "host_shadow", "shadow.cpp")
_HDR = os.path.join(os.path.dirname(_HERE), "solat_optics_b200", "csrc", "solat_math.cuh")
_SO = os.path.join(_HERE, "_build", "libsolat_shadow.so")
_lib = None
dp = ctypes.POINTER(ctypes.c_double)
ip = ctypes.POINTER(ctypes.c_int)
cd = ctypes.c_double
ci = ctypes.c_int


def lib():
    global _lib
    if _lib is None:
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        newest = max(os.path.getmtime(_SRC), os.path.getmtime(_HDR))
        if not os.path.exists(_SO) or os.path.getmtime(_SO) < newest:
            subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off",
                                   "-fvisibility=hidden", "-o", _SO, _SRC])
        _lib = ctypes.CDLL(_SO)
    return _lib


def D(a):
    a = np.ascontiguousarray(a, dtype=np.float64)
    return a, a.ctypes.data_as(dp)


def nf_source_rows(p, th, ph, n_si=3.4, n_vac=1.0):
    th, thp = D(th); ph, php = D(ph); p, pp = D(p)
    n = th.size * ph.size
    out = np.zeros((26, n)); st = np.zeros(n, dtype=np.int32)
    agg = lib().shadow_nf_source_rows(pp, thp, ci(th.size), php, ci(ph.size), cd(n_si), cd(n_vac),
                                      out.ctypes.data_as(dp), st.ctypes.data_as(ip))
    return out, st, agg


def nf_rx_rows(p, th, ph, th_fwhp, n_si=3.4, n_vac=1.0):
    th, thp = D(th); ph, php = D(ph); p, pp = D(p)
    T, P = np.meshgrid(th, ph)
    out = np.zeros((20, th.size * ph.size))
    agg = lib().shadow_nf_rx_rows(pp, thp, ci(th.size), php, ci(ph.size), cd(np.mean(T.ravel())), cd(np.mean(P.ravel())),
                                  cd(th_fwhp), cd(n_si), cd(n_vac), out.ctypes.data_as(dp))
    return out, agg


def nf_toy_rows(p, th, ph, y_ap, th_fwhp, n_si=3.4, n_vac=1.0):
    th, thp = D(th); ph, php = D(ph); p, pp = D(p)
    out = np.zeros((32, th.size * ph.size))
    agg = lib().shadow_nf_toy_rows(pp, thp, ci(th.size), php, ci(ph.size), cd(y_ap), cd(th_fwhp), cd(n_si), cd(n_vac),
                                   out.ctypes.data_as(dp))
    return out, agg


def nf_pixel(p, th, ph, rx_opl, rx_w, k, n_si=3.4, n_vac=1.0, packed=False):
    """packed=True traces two rays per step with the d2 packs, as the fused map kernel does."""
    th, thp = D(th); ph, php = D(ph); p, pp = D(p); ro, rop = D(rx_opl); rw, rwp = D(rx_w)
    o = np.zeros(2)
    fn = lib().shadow_nf_pixel2 if packed else lib().shadow_nf_pixel
    agg = fn(pp, thp, ci(th.size), php, ci(ph.size), cd(n_si), cd(n_vac), rop, rwp, cd(k),
                                o.ctypes.data_as(dp))
    return complex(o[0], o[1]), agg


def ff_mirror_rows(geo, th, ph, el, az):
    g, gp = D(geo); th, thp = D(th); ph, php = D(ph)
    out = np.zeros((9, th.size * ph.size))
    agg = lib().shadow_ff_mirror_rows(gp, thp, ci(th.size), php, ci(ph.size), cd(el), cd(az), out.ctypes.data_as(dp))
    return out, agg


def ff_field_rows(geo, th, ph, centre, el, az, pidx1=None, pidx2=None, pan1=None, pan2=None):
    g, gp = D(geo); th, thp = D(th); ph, php = D(ph); c, cp = D(centre)
    n = th.size * ph.size
    if pan1 is None:
        pan1 = np.zeros((10, 1)); pan1[0, 0] = pan1[1, 0] = 1; pan2 = pan1.copy()
        pidx1 = np.zeros(n, dtype=np.int32); pidx2 = np.zeros(n, dtype=np.int32)
    p1, p1p = D(pan1); p2, p2p = D(pan2)
    i1 = np.ascontiguousarray(pidx1, dtype=np.int32); i2 = np.ascontiguousarray(pidx2, dtype=np.int32)
    out = np.zeros((17, n))
    agg = lib().shadow_ff_field_rows(gp, thp, ci(th.size), php, ci(ph.size), cp, i1.ctypes.data_as(ip),
                                     i2.ctypes.data_as(ip), p1p, ci(p1.shape[1]), p2p, ci(p2.shape[1]), cd(el), cd(az),
                                     out.ctypes.data_as(dp))
    return out, agg


def ff_pixel(geo, th, ph, el, az, data_x, data_y, e_rec, k, packed=False):
    g, gp = D(geo); th, thp = D(th); ph, php = D(ph)
    gx, gxp = D(np.asarray(data_x)[0, :]); gy, gyp = D(data_y)
    e = np.ascontiguousarray(e_rec, dtype=np.complex128)
    o = np.zeros(2); nc = ci(0)
    fn = lib().shadow_ff_pixel2 if packed else lib().shadow_ff_pixel
    agg = fn(gp, thp, ci(th.size), php, ci(ph.size), cd(el), cd(az), gxp, gyp,
                                e.ctypes.data_as(dp), ci(gx.size), cd(k), o.ctypes.data_as(dp), ctypes.byref(nc))
    return complex(o[0], o[1]), nc.value, agg


def ff_row_intervals(geo, th, ph, el, az):
    """-> (rows whose interval differs from the per-ray test, rows that asked for the per-ray walk, (nph, 2) lo/hi)"""
    g, gp = D(geo); th, thp = D(th); ph, php = D(ph)
    lo_hi = np.zeros((ph.size, 2), dtype=np.int32); nf = ci(0)
    bad = lib().shadow_ff_row_intervals(gp, thp, ci(th.size), php, ci(ph.size), cd(el), cd(az),
                                        lo_hi.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), ctypes.byref(nf))
    return bad, nf.value, lo_hi


def ff_pred_check(geo, th, ph, el, az, gx_max, gy_max):
    """-> (accepted rays (negative: predictor off), worst predictor error over all rays [mm], worst sample error [mm],
    rays the predictor drops, wrongly dropped rays)"""
    g, gp = D(geo); th, thp = D(th); ph, php = D(ph)
    o = np.zeros(4)
    nv = lib().shadow_ff_pred_check(gp, thp, ci(th.size), php, ci(ph.size), cd(el), cd(az), cd(gx_max), cd(gy_max),
                                    o.ctypes.data_as(dp))
    return nv, o[0], o[1], int(o[2]), int(o[3])


def ff_keep_check(geo, th, ph, el, az, m, g_lo, g_hi, need_pos=True, need_neg=True, pile_min=-1):
    """-> (predictor on, accepted rays, kept rays, winner-table cells that differ between kept-only and all rays).
    pile_min >= 0: items that list at least that many rays also go through the scout stage (pile shrinking)."""
    g, gp = D(geo); th, thp = D(th); ph, php = D(ph)
    o = np.zeros(3)
    on = lib().shadow_ff_keep_check(gp, thp, ci(th.size), php, ci(ph.size), cd(el), cd(az), ci(m), cd(g_lo), cd(g_hi),
                                    ci(int(need_pos)), ci(int(need_neg)), o.ctypes.data_as(dp), ci(int(pile_min)))
    return on, int(o[0]), int(o[1]), int(o[2])


def regrid(field, n_new):
    f, fp = D(field)
    out = np.zeros((n_new, n_new))
    lib().shadow_regrid(fp, ci(f.shape[0]), ci(n_new), out.ctypes.data_as(dp))
    return out


def searchsorted_lean(axis, v):
    """(indices, axis_is_lean) of the kernel's branch-free searchsorted on a staged axis."""
    (a, ap), (v, vp) = D(axis), D(v)
    out = np.zeros(len(v), dtype=np.int32)
    lean = lib().shadow_searchsorted_lean(ap, ctypes.c_int(len(a)), vp, ctypes.c_int(len(v)),
                                          out.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    return out, bool(lean)


def cell_class(ux, uy, M, tol_x, tol_y):
    """ff_cell_class of the interpolated-cell kernel for predicted positions in cell units."""
    ux = np.ascontiguousarray(ux, dtype=np.float32); uy = np.ascontiguousarray(uy, dtype=np.float32)
    out = np.zeros(len(ux), dtype=np.int32)
    fp = ctypes.POINTER(ctypes.c_float)
    lib().shadow_cell_class(ux.ctypes.data_as(fp), uy.ctypes.data_as(fp), ctypes.c_int(len(ux)), ctypes.c_int(M),
                            ctypes.c_float(tol_x), ctypes.c_float(tol_y), out.ctypes.data_as(ctypes.POINTER(ctypes.c_int)))
    return out
