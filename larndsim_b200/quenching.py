"""Drop-in for ``larndsim.quenching`` (reference: larndsim/quenching.py:11-44)."""
import ctypes as C

from . import _launch as _l


@_l.kernel
def quench(tracks, mode):
    """``quench[BPG, TPB](tracks, mode)``: Box / Birks recombination, writes ``n_electrons`` and
    ``n_photons`` in place (quenching.py:23-44).  Raises ``ValueError`` for an invalid mode
    (quenching.py:37-38)."""
    c = _l.snapshot()
    if int(mode) not in (c.mode_box, c.mode_birks):
        raise ValueError("Invalid recombination mode: must be 'physics.BOX' or 'physics.BIRKS'")
    t = _l.dev(tracks, write=True, name="tracks", records=True)
    L = _l.layout(t)
    _l.check(_l.lib().lsb_quench(C.byref(c), C.byref(L), t.c, C.c_int64(t.shape[0]), C.c_int32(int(mode)), _l.stream()),
             "quench")
    _l.finish(t)
