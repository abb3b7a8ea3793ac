"""Drop-in for the device part of ``larndsim.fee`` (reference: larndsim/fee.py:499-655).
Packet export (fee.py:30-497) is host I/O and out of scope."""
import ctypes as C

import numpy as np
import torch

from . import _launch as _l
from . import rng as _rng


def digitize(integral_list, gain=None):
    """``digitize(integral_list[, gain])`` (fee.py:499-515): integrated charge -> ADC counts.
    Accepts and returns the caller's array kind (NumPy in -> NumPy out, device in -> torch CUDA out)."""
    c = _l.snapshot()
    host = isinstance(integral_list, np.ndarray)
    q = _l.dev(integral_list, want=np.float64, name="integral_list")
    n = q.size
    g = None
    if gain is not None and not np.isscalar(gain):
        g = _l.dev(gain, want=np.float64, name="gain")
        if g.size != n:
            raise ValueError("digitize: gain must have the shape of integral_list")
    elif gain is not None:
        gt = torch.full((n,), float(gain), dtype=torch.float64, device="cuda")
        g = _l.Dev(gt.data_ptr(), (n,), np.dtype("f8"), keep=gt)
    out = torch.empty(q.shape, dtype=torch.float64, device="cuda")
    _l.check(_l.lib().lsb_digitize(C.byref(c), q.c, g.c if g is not None else None, C.c_int64(n),
                                   C.c_void_p(out.data_ptr()), _l.stream()), "digitize")
    return out.cpu().numpy() if host else out


@_l.kernel
def get_adc_values(pixels_signals, pixels_signals_tracks, time_ticks, adc_list, adc_ticks_list, time_padding, rng_states,
                   current_fractions, pixel_thresholds):
    """``get_adc_values[BPG, TPB](pixels_signals, pixels_signals_tracks, time_ticks, adc_list,
    adc_ticks_list, time_padding, rng_states, current_fractions, pixel_thresholds)`` (fee.py:517-655)."""
    c = _l.snapshot()
    ps = _l.dev(pixels_signals, want=np.float64, name="pixels_signals")
    pst = _l.dev(pixels_signals_tracks, want=np.float64, name="pixels_signals_tracks")
    tt = _l.dev(time_ticks, want=np.float64, name="time_ticks")
    adc = _l.dev(adc_list, want=np.float64, write=True, name="adc_list")
    tks = _l.dev(adc_ticks_list, want=np.float64, write=True, name="adc_ticks_list")
    st, n_rng = _rng.states_dev(rng_states)
    cf = _l.dev(current_fractions, want=np.float64, write=True, name="current_fractions")
    thr = _l.dev(pixel_thresholds, want=np.float64, name="pixel_thresholds")
    U, Tt = ps.shape
    K = cf.shape[2]
    A = adc.shape[1]
    if pst.shape != (U, Tt, K) or tks.shape != adc.shape or cf.shape[:2] != (U, A) or thr.shape[0] < U:
        raise ValueError("get_adc_values: array shapes disagree")
    _l.check(_l.lib().lsb_get_adc_values(C.byref(c), ps.c, pst.c, C.c_int64(U), C.c_int32(Tt), C.c_int32(K), tt.c,
                                         C.c_int32(tt.shape[0]), adc.c, tks.c, C.c_int32(A), C.c_double(float(time_padding)),
                                         st.c, C.c_int64(n_rng), cf.c, thr.c, _l.stream()), "get_adc_values")
    _l.finish(adc, tks, cf, st)
