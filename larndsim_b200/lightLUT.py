"""Drop-in for ``larndsim.lightLUT`` (reference: larndsim/lightLUT.py:65-136)."""
import ctypes as C

import numpy as np

from . import _abi
from . import _launch as _l
from . import consts as _consts


def _lut_dev(lut):
    d = _l.dev(lut, name="lut", records=True)
    if len(d.shape) != 4 or d.dtype.fields is None:
        raise TypeError("light LUT must be a 4-D structured array [nx,ny,nz,ndet_tpc] with vis/t0/time_dist fields")
    return d, _abi.lut_layout(d.dtype, d.shape)


@_l.kernel
def calculate_light_incidence(tracks, lut, light_incidence, voxel):
    """``calculate_light_incidence[BPG, TPB](tracks, lut, light_incidence, voxel)``
    (lightLUT.py:65-136): visibility lookup -> ``n_photons_det`` / ``t0_det`` per optical channel."""
    c = _l.snapshot()
    t = _l.dev(tracks, name="tracks", records=True)
    L = _l.layout(t)
    ld, LL = _lut_dev(lut)
    li = _l.dev(light_incidence, write=True, name="light_incidence", records=True)
    LI = _abi.linc_layout(li.dtype)
    vx = _l.dev(voxel, want=np.int32, write=True, name="voxel")
    light = _consts.provider().light
    eff = _l.dev(np.ascontiguousarray(light.OP_CHANNEL_EFFICIENCY, dtype=np.float64), name="OP_CHANNEL_EFFICIENCY")
    tpc = _l.dev(np.ascontiguousarray(light.OP_CHANNEL_TO_TPC, dtype=np.int64), name="OP_CHANNEL_TO_TPC")
    S, ndet = li.shape
    _l.check(_l.lib().lsb_calculate_light_incidence(C.byref(c), C.byref(L), t.c, C.c_int64(S), ld.c, C.byref(LL), li.c,
                                                    C.byref(LI), C.c_int32(ndet), vx.c, eff.c, tpc.c, _l.stream()),
             "calculate_light_incidence")
    _l.finish(li, vx)
