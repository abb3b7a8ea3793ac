"""Drop-in for ``larndsim.pixels_from_track`` (reference: larndsim/pixels_from_track.py)."""
import ctypes as C

import numpy as np

from . import _launch as _l
from . import consts as _consts

MAX_NEIGHBOR_BACKTRACK_DISTANCE = _consts.MAX_NEIGHBOR_BACKTRACK_DISTANCE


def pixel2id(pixel_x, pixel_y, pixel_plane):
    """x,y,plane -> unique id (pixels_from_track.py:13-26); host helper (also used by fee.py:147)."""
    n = _consts.provider().detector.N_PIXELS
    return pixel_x + n[0] * (pixel_y + n[1] * pixel_plane)


def id2pixel(pid):
    """unique id -> (x, y, plane) with Python floor semantics (pixels_from_track.py:28-41)."""
    n = _consts.provider().detector.N_PIXELS
    return (pid % n[0], (pid // n[0]) % n[1], pid // (n[0] * n[1]))


@_l.kernel
def max_pixels(tracks, n_max_pixels):
    """``max_pixels[BPG, TPB](tracks, n_max_pixels)``: atomic max of the Bresenham step count into
    ``n_max_pixels[0]`` (pixels_from_track.py:43-65)."""
    c = _l.snapshot()
    t = _l.dev(tracks, name="tracks", records=True)
    L = _l.layout(t)
    m = _l.dev(n_max_pixels, want=np.int64, write=True, name="n_max_pixels")
    _l.check(_l.lib().lsb_max_pixels(C.byref(c), C.byref(L), t.c, C.c_int64(t.shape[0]), m.c, _l.stream()), "max_pixels")
    _l.finish(m)


@_l.kernel
def get_pixels(tracks, active_pixels, neighboring_pixels, neighboring_radius, n_pixels_list, radius):
    """``get_pixels[BPG, TPB](tracks, active_pixels, neighboring_pixels, neighboring_radius,
    n_pixels_list, radius)`` (pixels_from_track.py:67-109).  Outputs are caller-initialised (-1)."""
    c = _l.snapshot()
    t = _l.dev(tracks, name="tracks", records=True)
    L = _l.layout(t)
    a = _l.dev(active_pixels, want=np.int32, write=True, name="active_pixels")
    nb = _l.dev(neighboring_pixels, want=np.int32, write=True, name="neighboring_pixels")
    nr = _l.dev(neighboring_radius, want=np.int32, write=True, name="neighboring_radius")
    npl = _l.dev(n_pixels_list, want=np.float64, write=True, name="n_pixels_list")
    S = t.shape[0]
    if a.shape[0] != S or nb.shape[0] != S or nr.shape != nb.shape or npl.shape[0] != S:
        raise ValueError("get_pixels: output shapes do not match the number of tracks")
    _l.check(_l.lib().lsb_get_pixels(C.byref(c), C.byref(L), t.c, C.c_int64(S), a.c, C.c_int32(a.shape[1]), nb.c, nr.c,
                                     C.c_int32(nb.shape[1]), npl.c, C.c_int32(int(radius)), _l.stream()), "get_pixels")
    _l.finish(a, nb, nr, npl)
