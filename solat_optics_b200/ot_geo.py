"""Optics-tube (LATR) geometry: drop-in for the reference's `solat_optics/ot_geo.py`.

Same public names, argument meaning and numbers (reference ot_geo.py:12-71 for LatGeo and the
offsets, :74-139,186-250,297-405 for sags and slopes, :142-183,253-294,408-531 for the frames).
These are small host-side geometry definitions (plot helpers, user scripts); the ray tracing
itself runs in the CUDA library, which carries the same constants (csrc/solat_math.cuh).

Out of scope: `plot_lenses` and `filter_geo_new` (matplotlib / a PNG asset the
reference does not ship); they raise NotImplementedError here.
"""
import numpy as np


class LatGeo:
    """LAT receiver optics-tube settings (reference ot_geo.py:12-58)."""

    def __init__(self):
        self.n_si = 3.4
        self.n_vac = 1.0
        self.th_fwhp = 35 * np.pi / 180
        self.n_scan = 100
        self.de_ang = 0.5 / 60 * np.pi / 180
        self.lambda_ = (30.0 / 150.0) * 0.01  # [m]
        self.k = 2 * np.pi / self.lambda_
        self.lyot_y = 453.62
        self.x_ap = 0
        self.y_ap = 0
        self.z_ap = 0

    def set_fwhp(self, fwhp):
        self.th_fwhp = fwhp

    def set_nscan(self, n_val):
        self.n_scan = n_val

    def set_res(self, res):
        self.de_ang = res

    def set_wavelength(self, wavelength):
        self.lambda_ = wavelength
        self.k = 2 * np.pi / wavelength


# tilt angles and axial offsets, reference ot_geo.py:61-71
th1_l1 = np.pi / 2
th2_l1 = -np.pi / 2
th1_l2 = -np.pi / 2
th2_l2 = -np.pi / 2
th1_l3 = -np.pi / 2
th2_l3 = -np.pi / 2

LENS1_Y = 975.08
LENS2_Y = 592.76
LENS3_Y = 180.8 - ((2.5 + 1.513) - (35.16 - 32.19))
LYOT_Y = 453.62


class _EvenAsphere:
    """z = c r^2/(1+sqrt(1-(1+k)c^2 r^2)) + a1 r^2 + a2 r^4 + a3 r^6 + a4 r^8"""

    def __init__(self, radius, conic, a_1, a_2, a_3, a_4):
        self.c = 1 / radius
        self.k = conic
        self.a = (a_1, a_2, a_3, a_4)

    def _root(self, r):
        return np.sqrt(1 - ((1 + self.k) * self.c ** 2 * r ** 2))

    def sag(self, x_arr, y_arr):
        r = np.sqrt(x_arr ** 2 + y_arr ** 2)
        a_1, a_2, a_3, a_4 = self.a
        z = (self.c * r ** 2) / (1 + self._root(r))
        return z + a_1 * r ** 2 + a_2 * r ** 4 + a_3 * r ** 6 + a_4 * r ** 8

    def slope(self, x_arr, y_arr):
        r = np.sqrt(x_arr ** 2 + y_arr ** 2)
        a_1, a_2, a_3, a_4 = self.a
        s = self._root(r)
        radial = (a_1 * 2) + (a_2 * 4 * r ** 2) + (a_3 * 6 * r ** 4) + (a_4 * 8 * r ** 6)
        radial = radial + (self.c * 2) / (1 + s) + (self.c ** 3 * (self.k + 1) * r ** 2) / (s * (1 + s) ** 2)
        return x_arr * radial, y_arr * radial


_L1A = _EvenAsphere(-1510.61, 12.0966, -5.876941e-5, 1.769198e-9, -1.310183e-15, 1.994766e-19)
_L2A = _EvenAsphere(-871.25, -10.201256, -1.147640e-04, 1.103643e-10, -2.969608e-14, 5.980271e-19)
_L3A = _EvenAsphere(6608.30, 414.142695, -2.981594e-06, 3.955202e-10, 3.180727e-16, 0)
_L3B = _EvenAsphere(604.44, -3.177341, 2.366829e-05, -1.000417e-10, -6.565681e-15, 0)


def z1b():
    return 0


def z2b():
    return 0


def z_lyot():
    return 0


def filt():
    return 0


def d_z1b():
    return 0, 0


def d_z2b():
    return 0, 0


def d_zlyot():
    return 0, 0


def z1a(x_arr, y_arr):
    return _L1A.sag(x_arr, y_arr)


def z2a(x_arr, y_arr):
    return _L2A.sag(x_arr, y_arr)


def z3a(x_arr, y_arr):
    return _L3A.sag(x_arr, y_arr)


def z3b(x_arr, y_arr):
    return _L3B.sag(x_arr, y_arr)


def d_z1a(x_arr, y_arr):
    return _L1A.slope(x_arr, y_arr)


def d_z2a(x_arr, y_arr):
    return _L2A.slope(x_arr, y_arr)


def d_z3a(x_arr, y_arr):
    return _L3A.slope(x_arr, y_arr)


def d_z3b(x_arr, y_arr):
    return _L3B.slope(x_arr, y_arr)


class _Frame:
    """Surface frame: shift along the tube axis by `offset`, tilt about x by `tilt`."""

    def __init__(self, offset, tilt):
        self.offset = offset
        self.tilt = tilt

    def from_tele(self, x_arr, y_arr, z_arr):
        y_s = y_arr + self.offset
        c, s = np.cos(-self.tilt), np.sin(-self.tilt)
        return x_arr, y_s * c - z_arr * s, y_s * s + z_arr * c

    def to_tele(self, x_arr, y_arr, z_arr):
        c, s = np.cos(self.tilt), np.sin(self.tilt)
        return x_arr, (y_arr * c - z_arr * s) - self.offset, y_arr * s + z_arr * c


_F1B = _Frame(LENS1_Y, th1_l1)
_F1A = _Frame(LENS1_Y - 19.32, th2_l1)
_F2B = _Frame(LENS2_Y, th1_l2)
_F2A = _Frame(LENS2_Y - 26.47, th2_l2)
_F2A_IN = _Frame(LENS2_Y - 26.47, th2_l1)   # the reference's tele_into_m2a tilts by th2_l1 (ot_geo.py:291)
_F3A = _Frame(LENS3_Y, th1_l3)
_F3B = _Frame(LENS3_Y + 35.16, th2_l3)
_FLYOT = _Frame(LYOT_Y, th1_l3)
_FFP = _Frame(0.0, th2_l3)
_FF1 = _Frame(991.84174, th1_l3)
_FF2 = _Frame(608.56983, th1_l3)
_FF3 = _Frame(18.10815, th1_l3)


def m1b_into_tele(x_arr, y_arr, z_arr):
    return _F1B.to_tele(x_arr, y_arr, z_arr)


def m1a_into_tele(x_arr, y_arr, z_arr):
    return _F1A.to_tele(x_arr, y_arr, z_arr)


def tele_into_m1b(x_arr, y_arr, z_arr):
    return _F1B.from_tele(x_arr, y_arr, z_arr)


def tele_into_m1a(x_arr, y_arr, z_arr):
    return _F1A.from_tele(x_arr, y_arr, z_arr)


def m2b_into_tele(x_arr, y_arr, z_arr):
    return _F2B.to_tele(x_arr, y_arr, z_arr)


def m2a_into_tele(x_arr, y_arr, z_arr):
    return _F2A.to_tele(x_arr, y_arr, z_arr)


def tele_into_m2b(x_arr, y_arr, z_arr):
    return _F2B.from_tele(x_arr, y_arr, z_arr)


def tele_into_m2a(x_arr, y_arr, z_arr):
    return _F2A_IN.from_tele(x_arr, y_arr, z_arr)


def m3a_into_tele(x_arr, y_arr, z_arr):
    return _F3A.to_tele(x_arr, y_arr, z_arr)


def m3b_into_tele(x_arr, y_arr, z_arr):
    return _F3B.to_tele(x_arr, y_arr, z_arr)


def fp_into_tele(x_arr, y_arr, z_arr):
    return _FFP.to_tele(x_arr, y_arr, z_arr)


def tele_into_f1(x_arr, y_arr, z_arr):
    return _FF1.from_tele(x_arr, y_arr, z_arr)


def tele_into_f2(x_arr, y_arr, z_arr):
    return _FF2.from_tele(x_arr, y_arr, z_arr)


def tele_into_f3(x_arr, y_arr, z_arr):
    return _FF3.from_tele(x_arr, y_arr, z_arr)


def tele_into_m3a(x_arr, y_arr, z_arr):
    return _F3A.from_tele(x_arr, y_arr, z_arr)


def tele_into_m3b(x_arr, y_arr, z_arr):
    return _F3B.from_tele(x_arr, y_arr, z_arr)


def lyot_into_tele(x_arr, y_arr, z_arr):
    return _FLYOT.to_tele(x_arr, y_arr, z_arr)


def tele_into_lyot(x_arr, y_arr, z_arr):
    return _FLYOT.from_tele(x_arr, y_arr, z_arr)


def plot_lenses():
    raise NotImplementedError("plot_lenses is a matplotlib helper outside the beam-simulation hot path")


def filter_geo_new():
    raise NotImplementedError("filter_geo_new needs filter/so_filter2.png, which the reference does not ship")
