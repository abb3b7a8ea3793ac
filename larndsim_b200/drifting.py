"""Drop-in for ``larndsim.drifting`` (reference: larndsim/drifting.py:11-58)."""
import ctypes as C

from . import _launch as _l


@_l.kernel
def drift(tracks):
    """``drift[BPG, TPB](tracks)``: TPC lookup, lifetime attenuation, diffusion widths and arrival
    times, in place (drifting.py:26-58)."""
    c = _l.snapshot()
    t = _l.dev(tracks, write=True, name="tracks", records=True)
    L = _l.layout(t)
    _l.check(_l.lib().lsb_drift(C.byref(c), C.byref(L), t.c, C.c_int64(t.shape[0]), _l.stream()), "drift")
    _l.finish(t)
