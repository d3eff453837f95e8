"""Phase utilities used by the far-field CLI: drop-in for `do_unwrap` / `twodunwrap` /
`twodunwrapx1` of the reference's `solat_optics/phase_correct.py` (:8-92).

One-shot O(N^2) host preprocessing of the measured near-field map (scripts/ff_sim.py:84); it is not
per-pixel work and stays on the host.  The measurement-fitting helpers of
the reference module (defocus, beam_centering, gauss...) are outside the simulation path and are
not provided.
"""
import numpy as np

_N_OFF = 1000  # candidate offsets are (-1000 .. 999) * 2 pi, as in the reference


def _snap(nxt, cur):
    """For every element pick m in [-1000, 999] minimising |nxt + 2 pi m - cur| (first minimum wins)
    and return nxt + 2 pi m, evaluated with the reference's own expression for the offset."""
    guess = np.rint((cur - nxt) / (2 * np.pi))
    best_val = None
    best_diff = None
    for dm in (-1, 0, 1):          # ascending m, strict '<' keeps the first minimum like np.where(...)[0]
        m = np.clip(guess + dm, -_N_OFF, _N_OFF - 1)
        cand = nxt + m * 2 * np.pi
        diff = np.abs(cand - cur)
        if best_val is None:
            best_val, best_diff = cand, diff
        else:
            take = diff < best_diff
            best_val = np.where(take, cand, best_val)
            best_diff = np.where(take, diff, best_diff)
    return best_val


def twodunwrapx1(array):
    """Unwrap each row outward from the centre column, in the reference's four sweeps (in place)."""
    end = len(array)
    ncol = np.shape(array)[1]
    mid = int(end / 2)
    sweeps = (
        (range(mid, end), range(mid, ncol - 1), +1),
        (range(mid, 0, -1), range(mid, 0, -1), -1),
        (range(mid, end - 1), range(mid, 0, -1), -1),
        (range(mid, 0, -1), range(mid, end - 1), +1),
    )
    for rows, cols, step in sweeps:
        rows = np.fromiter(rows, dtype=int)
        if rows.size == 0:
            continue
        for j in cols:
            array[rows, j + step] = _snap(array[rows, j + step], array[rows, j])
    return array


def twodunwrap(array):
    xunwraped = twodunwrapx1(np.transpose(array))
    return twodunwrapx1(np.transpose(xunwraped))


def do_unwrap(phase_input):
    """Unwrap a 2-D phase map given in DEGREES; returns degrees with the [0, 0] offset removed."""
    phi = phase_input * (np.pi / 180)
    unwrapped = twodunwrap(phi)
    unwrapped = unwrapped - unwrapped[0, 0]
    return unwrapped * (180 / np.pi)
