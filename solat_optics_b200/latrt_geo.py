"""LAT mirror geometry: drop-in for the reference's `solat_optics/latrt_geo.py`.

Same public names and numbers (reference latrt_geo.py:7-49 geometry bag, :54-73 polynomial
coefficients, :77-150 sag/slope, :153-294 frames).  Host-side definitions; the kernels carry
the same coefficients in Horner form (csrc/solat_math.cuh m1_sag / m2_sag).
"""
import numpy as np


class initialize_telescope_geometry:
    """Telescope parameters (class attributes, instantiated and mutated by callers)."""

    F_2 = 7000
    th_1 = np.arctan(1 / 2)          # primary tilt
    th_2 = np.arctan(1 / 3)          # secondary tilt
    th2 = (-np.pi / 2) - th_2
    th_fwhp = 35 * np.pi / 180
    N_scan = 100
    de_ang = 2 / 60 * np.pi / 180
    lambda_ = (30.0 / 150.0) * 0.01  # [m]
    k = 2 * np.pi / lambda_

    rx_x = 0
    rx_y = 0
    rx_z = 0

    x_phref = 0
    y_phref = -7.2
    z_phref = 0

    x_rotc = 0
    y_rotc = -7.2
    z_rotc = 0

    x_tow = 0
    y_tow = -7.2
    z_tow = 1e3

    az0 = 0
    el0 = 0

    x_ap = 3.0
    y_ap = -7.2
    z_ap = 4.0


# 2-D polynomial coefficients a[i, j] of (x/R_N)^i (y/R_N)^j  [mm]
a1 = np.zeros((7, 7))
a1[0, 2:] = [-57.74022, 1.5373825, 1.154294, -0.441762, 0.0906601]
a1[2, :5] = [-72.17349, 1.8691899, 2.8859421, -1.026471, 0.2610568]
a1[4, :3] = [1.8083973, -0.603195, 0.2177414]
a1[6, 0] = 0.0394559

a2 = np.zeros((8, 8))
a2[0, 2:] = [103.90461, 6.6513025, 2.8405781, -0.7819705, -0.0400483, 0.0896645]
a2[2, :6] = [115.44758, 7.3024355, 5.7640389, -1.578144, -0.0354326, 0.2781226]
a2[4, :4] = [2.9130983, -0.8104051, -0.0185283, 0.2626023]
a2[6, :2] = [-0.0250794, 0.0709672]

R_N = 3000  # [mm]


def _poly(coef, x_arr, y_arr):
    amp = 0
    for (i, j), c in np.ndenumerate(coef):
        amp += c * ((x_arr / R_N) ** i) * ((y_arr / R_N) ** j)
    return amp


def _poly_grad(coef, x_arr, y_arr):
    amp_x = 0
    amp_y = 0
    for (i, j), c in np.ndenumerate(coef):
        amp_x += c * (i / R_N) * ((x_arr / R_N) ** (i - 1)) * ((y_arr / R_N) ** j)
        amp_y += c * ((x_arr / R_N) ** i) * (j / R_N) * ((y_arr / R_N) ** (j - 1))
    return amp_x, amp_y


def z1(x_arr, y_arr):
    return _poly(a1, x_arr, y_arr)


def z2(x_arr, y_arr):
    return _poly(a2, x_arr, y_arr)


def d_z1(x_arr, y_arr):
    return _poly_grad(a1, x_arr, y_arr)


def d_z2(x_arr, y_arr):
    return _poly_grad(a2, x_arr, y_arr)


def _rot_x(y, z, ang):
    return y * np.cos(ang) - z * np.sin(ang), y * np.sin(ang) + z * np.cos(ang)


def rotate_az_el(x_arr, y_arr, z_arr, el_, az_):
    """Rotate about the rotation centre by elevation then azimuth (reference :272-294)."""
    g = initialize_telescope_geometry
    cx, cy, cz = g.x_rotc * 1e3, g.y_rotc * 1e3, g.z_rotc * 1e3
    x0, y0, z0 = x_arr - cx, y_arr - cy, z_arr - cz
    y1, z1_ = _rot_x(y0, z0, el_)
    x2 = np.cos(az_) * x0 + np.sin(az_) * z1_
    z2_ = -np.sin(az_) * x0 + np.cos(az_) * z1_
    return x2 + cx, y1 + cy, z2_ + cz


def m1_into_tele(x_arr, y_arr, z_arr):
    th1 = initialize_telescope_geometry.th_1
    x_t = x_arr * np.cos(np.pi) + z_arr * np.sin(np.pi)
    z_t = -x_arr * np.sin(np.pi) + z_arr * np.cos(np.pi)
    y_r, z_r = _rot_x(y_arr, z_t, th1)
    return x_t, y_r - 7200, z_r - 3600


def m2_into_tele(x_arr, y_arr, z_arr):
    th2 = initialize_telescope_geometry.th2
    x_t = x_arr * np.cos(np.pi) - y_arr * np.sin(np.pi)
    y_t = x_arr * np.sin(np.pi) + y_arr * np.cos(np.pi)
    y_r, z_r = _rot_x(y_t, z_arr, th2)
    return x_t, y_r - 4800 - 7200, z_r


def tele_into_m1(x_arr, y_arr, z_arr, el_, az_):
    x_r, y_r, z_r = rotate_az_el(x_arr, y_arr, z_arr, el_, az_)
    th1 = initialize_telescope_geometry.th_1
    y_t, z_t = _rot_x(y_r + 7200, z_r + 3600, -th1)
    return (x_r * np.cos(np.pi) + z_t * np.sin(np.pi), y_t, -x_r * np.sin(np.pi) + z_t * np.cos(np.pi))


def tele_into_m2(x_arr, y_arr, z_arr, el_, az_):
    x_r, y_r, z_r = rotate_az_el(x_arr, y_arr, z_arr, el_, az_)
    th2 = initialize_telescope_geometry.th2
    y_t, z_t = _rot_x(y_r + 4800 + 7200, z_r, -th2)
    return (x_r * np.cos(-np.pi) - y_t * np.sin(-np.pi), x_r * np.sin(-np.pi) + y_t * np.cos(-np.pi), z_t)


def z_ap(x_arr, y_arr):
    return 0 * x_arr * y_arr


def z_focal(x_arr, y_arr):
    return 0 * x_arr * y_arr


def tele_into_apert(x_arr, y_arr, z_arr, el_, az_):
    x_r, y_r, z_r = rotate_az_el(x_arr, y_arr, z_arr, el_, az_)
    return x_r, y_r, z_r - 4e3


def foc_into_tele(x_arr, y_arr, z_arr, el_, az_):
    y_t, z_t = _rot_x(y_arr, z_arr, np.pi / 2)
    return rotate_az_el(x_arr, y_t, z_t, el_, az_)


def tele_into_foc(x_arr, y_arr, z_arr, el_, az_):
    x_r, y_r, z_r = rotate_az_el(x_arr, y_arr, z_arr, el_, az_)
    y_t, z_t = _rot_x(y_r, z_r, -np.pi / 2)
    return x_r, y_t, z_t
