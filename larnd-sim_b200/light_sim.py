"""Drop-in for the waveform kernels of ``larndsim.light_sim`` (reference: larndsim/light_sim.py:58-336).
Trigger search, noise synthesis, digitisation and export (:339-780) are downstream of the path and
out of scope."""
import ctypes as C

import numpy as np

from . import _abi
from . import _launch as _l
from . import consts as _consts
from . import rng as _rng
from .lightLUT import _lut_dev


def _truth(ids, photons, write, what):
    i = _l.dev(ids, want=np.int64, write=write, name=what + "_true_track_id")
    p = _l.dev(photons, want=np.float64, write=write, name=what + "_true_photons")
    n = i.shape[-1] if len(i.shape) == 3 else 0
    return i, p, n


@_l.kernel
def sum_light_signals(segments, segment_voxel, segment_track_id, light_inc, op_channel, lut, start_time, light_sample_inc,
                      light_sample_inc_true_track_id, light_sample_inc_true_photons, sorted_indices, t0_profile_length):
    """``sum_light_signals[BPG, TPB](...)`` (light_sim.py:58-129): photons per (channel, tick)."""
    c = _l.snapshot()
    sg = _l.dev(segments, name="segments", records=True)
    L = _l.layout(sg)
    vx = _l.dev(segment_voxel, want=np.int32, name="segment_voxel")
    tid = _l.dev(segment_track_id, want=np.int64, name="segment_track_id")
    li = _l.dev(light_inc, name="light_inc", records=True)
    LI = _abi.linc_layout(li.dtype)
    oc = _l.dev(op_channel, want=np.int32, name="op_channel")
    ld, LL = _lut_dev(lut)
    out = _l.dev(light_sample_inc, want=np.float32, write=True, name="light_sample_inc")
    ti, tp, n_true = _truth(light_sample_inc_true_track_id, light_sample_inc_true_photons, True, "light_sample_inc")
    si = _l.dev(sorted_indices, want=np.int64, name="sorted_indices")
    ndet, nticks = out.shape
    n_sorted = si.shape[1] if len(si.shape) == 2 else 0
    _l.check(_l.lib().lsb_sum_light_signals(C.byref(c), C.byref(L), sg.c, C.c_int64(sg.shape[0]), vx.c, tid.c, li.c,
                                            C.byref(LI), C.c_int32(li.shape[1]), oc.c, ld.c, C.byref(LL),
                                            C.c_double(float(start_time)), out.c, C.c_int32(ndet), C.c_int32(nticks),
                                            ti.c, tp.c, C.c_int32(n_true), si.c, C.c_int64(n_sorted),
                                            C.c_double(float(t0_profile_length)), _l.stream()), "sum_light_signals")
    _l.finish(out, ti, tp)


@_l.kernel
def calc_scintillation_effect(light_sample_inc, light_sample_inc_true_track_id, light_sample_inc_true_photons,
                              light_sample_inc_scint, light_sample_inc_scint_true_track_id,
                              light_sample_inc_scint_true_photons):
    """``calc_scintillation_effect[BPG, TPB](6 arrays)`` (light_sim.py:148-183): causal FIR with the
    singlet/triplet scintillation time profile."""
    c = _l.snapshot()
    a = _l.dev(light_sample_inc, want=np.float32, name="light_sample_inc")
    ai, ap, n_in = _truth(light_sample_inc_true_track_id, light_sample_inc_true_photons, False, "light_sample_inc")
    o = _l.dev(light_sample_inc_scint, want=np.float32, write=True, name="light_sample_inc_scint")
    oi, op, n_out = _truth(light_sample_inc_scint_true_track_id, light_sample_inc_scint_true_photons, True, "light_sample_inc_scint")
    ndet, nticks = a.shape
    _l.check(_l.lib().lsb_calc_scintillation_effect(C.byref(c), a.c, ai.c, ap.c, o.c, oi.c, op.c, C.c_int32(ndet),
                                                    C.c_int32(nticks), C.c_int32(n_in), C.c_int32(n_out), _l.stream()),
             "calc_scintillation_effect")
    _l.finish(o, oi, op)


@_l.kernel
def calc_stat_fluctuations(light_sample_inc, light_sample_inc_disc, rng_states):
    """``calc_stat_fluctuations[BPG, TPB](in, out, rng_states)`` (light_sim.py:220-238)."""
    c = _l.snapshot()
    a = _l.dev(light_sample_inc, want=np.float32, name="light_sample_inc")
    o = _l.dev(light_sample_inc_disc, want=np.float32, write=True, name="light_sample_inc_disc")
    st, n_rng = _rng.states_dev(rng_states)
    ndet, nticks = a.shape
    _l.check(_l.lib().lsb_calc_stat_fluctuations(C.byref(c), a.c, o.c, C.c_int32(ndet), C.c_int32(nticks), st.c,
                                                 C.c_int64(n_rng), _l.stream()), "calc_stat_fluctuations")
    _l.finish(o, st)


@_l.kernel
def calc_light_detector_response(light_sample_inc, light_sample_inc_true_track_id, light_sample_inc_true_photons,
                                 light_response, light_response_true_track_id, light_response_true_photons):
    """``calc_light_detector_response[BPG, TPB](6 arrays)`` (light_sim.py:303-336): causal FIR with the
    SiPM impulse response times ``light.LIGHT_GAIN``."""
    c = _l.snapshot()
    light = _consts.provider().light
    a = _l.dev(light_sample_inc, want=np.float32, name="light_sample_inc")
    ai, ap, n_in = _truth(light_sample_inc_true_track_id, light_sample_inc_true_photons, False, "light_sample_inc")
    o = _l.dev(light_response, want=np.float32, write=True, name="light_response")
    oi, op, n_out = _truth(light_response_true_track_id, light_response_true_photons, True, "light_response")
    ndet, nticks = a.shape
    gain = np.ascontiguousarray(np.asarray(light.LIGHT_GAIN, dtype=np.float64).reshape(-1))
    if gain.size < ndet:
        raise ValueError("light.LIGHT_GAIN has fewer entries than light_sample_inc has channels")
    g = _l.dev(gain, name="LIGHT_GAIN")
    imp = getattr(light, "IMPULSE_MODEL", None)
    if imp is not None:
        # the tap weights are evaluated on the host once per call: IMPULSE_MODEL stays a host array
        imph = np.ascontiguousarray(imp, dtype=np.float64)
        imp_c, n_imp = imph.ctypes.data_as(C.c_void_p), imph.size
    else:
        imph, imp_c, n_imp = None, None, 0
    _l.check(_l.lib().lsb_calc_light_detector_response(C.byref(c), a.c, ai.c, ap.c, o.c, oi.c, op.c, C.c_int32(ndet),
                                                       C.c_int32(nticks), C.c_int32(n_in), C.c_int32(n_out), g.c, imp_c,
                                                       C.c_int32(n_imp), _l.stream()), "calc_light_detector_response")
    _l.finish(o, oi, op)
