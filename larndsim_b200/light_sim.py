"""Drop-in for ``larndsim.light_sim``: the waveform kernels (reference: larndsim/light_sim.py:58-336) and the trigger
search + digitisation stage (``get_triggers`` :380-477, ``sim_triggers`` :545-619, ``digitize_signal`` :480-543,
``gen_light_detector_noise`` :339-377; CUDA: csrc/light_trigger.cuh).  HDF5 export (:621-780) is host I/O and not
rebuilt."""
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


# ---------------------------------------------------------------------------------------------------------
# extent of the light window (the callers in front of sum_light_signals, cli/simulate_pixels.py:1119-1135)
# ---------------------------------------------------------------------------------------------------------
def _extent(light_incidence):
    import torch
    from . import _abi
    li = _l.dev(light_incidence, name="light_incidence", records=True)
    if len(li.shape) != 2:
        raise ValueError("light_incidence must have shape (ntracks, ndet)")
    S, ndet = li.shape
    LI = _abi.linc_layout(li.dtype)
    mm = torch.empty(2, dtype=torch.float32, device="cuda")
    active = torch.empty(ndet, dtype=torch.uint8, device="cuda")
    _l.check(_l.lib().lsb_light_extent(li.c, C.byref(LI), C.c_int64(S), C.c_int32(ndet), C.c_void_p(mm.data_ptr()),
                                       C.c_void_p(active.data_ptr()), _l.stream()), "light_extent")
    return mm, active


def get_nticks(light_incidence):
    """``get_nticks(light_incidence)`` (light_sim.py:24-42) -> (number of light ticks, time of the first tick [us]).
    The arithmetic on the float32 extremes is NumPy's (float32 scalars against Python floats), as in the reference."""
    light = _consts.provider().light
    mm, active = _extent(light_incidence)
    lo, hi = (np.float32(v) for v in mm.cpu().numpy())
    # the reference's constants are Python numbers (consts/light.py:122-123): weakly typed against the float32 extremes
    w0, w1, tick = (v.item() if hasattr(v, "item") else v for v in (light.LIGHT_WINDOW[0], light.LIGHT_WINDOW[1], light.LIGHT_TICK_SIZE))
    if bool(active.any().item()) and light.LIGHT_TRIG_MODE == 0:
        start_time = lo - w0
        end_time = hi + w1
        return int(np.ceil((end_time - start_time) / tick)), start_time
    return int((w1 + w0) / tick), 0


def get_active_op_channel(light_incidence):
    """``get_active_op_channel(light_incidence)`` (light_sim.py:44-57) -> int32 CUDA tensor of the channels some segment
    gives photons to."""
    import torch
    _, active = _extent(light_incidence)
    return torch.nonzero(active).reshape(-1).to(torch.int32)


# ---------------------------------------------------------------------------------------------------------
# trigger search and digitisation
# ---------------------------------------------------------------------------------------------------------
def _host_ids(a):
    import torch
    if isinstance(a, np.ndarray):
        return a
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    if hasattr(a, "get"):
        return np.asarray(a.get())
    if hasattr(a, "__cuda_array_interface__"):
        return torch.as_tensor(a, device="cuda").cpu().numpy()
    return np.asarray(a)


def _wave(signal, name="signal"):
    """waveform array as a device buffer in its own precision (float32 or float64)"""
    import torch
    dt = np.dtype(str(signal.dtype).replace("torch.", "")) if not isinstance(signal, np.ndarray) else signal.dtype
    want = np.float64 if dt == np.dtype("f8") else np.float32
    return _l.dev(signal, want=want, name=name), (1 if want is np.float64 else 0)


def _tpc_to_module(d):
    if hasattr(d, "TPC_TO_MODULE"):
        return {int(k): int(v) for k, v in dict(d.TPC_TO_MODULE).items()}
    return {int(t): int(m) for m, tpcs in dict(d.MODULE_TO_TPCS).items() for t in np.asarray(tpcs).ravel()}


def get_triggers(signal, group_threshold, op_channel_idx, i_subbatch):
    """``get_triggers(signal, group_threshold, op_channel_idx, i_subbatch)`` (light_sim.py:380-477) ->
    ``(trigger tick indices [ntrigs], op channel indices [ntrigs, ndet_module], trigger types [ntrigs])`` as host arrays.
    The group sums, block averages, threshold test and the sequential per-module search run on the GPU."""
    import torch
    from math import ceil
    p = _consts.provider()
    li, d = p.light, p.detector
    op = np.asarray(_host_ids(op_channel_idx)).astype(np.int64)
    if li.LIGHT_TRIG_MODE == 1:
        if i_subbatch == 0:
            return np.array([0]), np.array([op]), np.array([1])
        return np.empty((0,), dtype=int), np.empty((0, len(op)), dtype=int), np.empty((0,), dtype=int)
    if li.LIGHT_TRIG_MODE != 0:
        return np.empty((0,), dtype=int), np.empty((0, len(op)), dtype=int), np.empty((0,), dtype=int)
    sig, f64 = _wave(signal)
    ndet, nticks = sig.shape
    cpt = int(li.OP_CHANNEL_PER_TRIG)
    sf = round(li.LIGHT_DIGIT_SAMPLE_SPACING / li.LIGHT_TICK_SIZE)
    digit_ticks = ceil((li.LIGHT_TRIG_WINDOW[1] + li.LIGHT_TRIG_WINDOW[0]) / li.LIGHT_TICK_SIZE)
    # the modules the simulated channels belong to, and each channel's module slot (light_sim.py:414-429)
    t2m = _tpc_to_module(d)
    m2t = {int(k): [int(x) for x in np.asarray(v).ravel()] for k, v in dict(d.MODULE_TO_TPCS).items()}
    tpc_ids = np.unique(np.asarray(li.OP_CHANNEL_TO_TPC)[op])
    mod_ids = np.unique([t2m[int(t)] for t in tpc_ids])
    t2c = np.asarray(li.TPC_TO_OP_CHANNEL)
    mod_channels = [t2c[m2t[int(m)]].ravel() for m in mod_ids]
    chan_module = np.full(ndet, -1, dtype=np.int32)
    for slot in reversed(range(len(mod_ids))):                  # a channel listed by several modules counts for each: not possible
        chan_module[np.isin(op, mod_channels[slot])] = slot      # with disjoint channel lists (asserted below)
    assert sum(int(np.isin(op, mc).sum()) for mc in mod_channels) == int((chan_module >= 0).sum()), "op channels shared between modules"
    thr = _l.dev(np.ascontiguousarray(_host_ids(group_threshold), dtype=np.float64), want=np.float64)
    cm = _l.dev(chan_module, want=np.int32)
    nmod = len(mod_ids)
    max_trig = int(nticks // max(digit_ticks, 1)) + 2
    tidx = torch.zeros((max(nmod, 1), max_trig), dtype=torch.int64, device="cuda")
    ntr = torch.zeros(max(nmod, 1), dtype=torch.int32, device="cuda")
    _l.check(_l.lib().lsb_light_get_triggers(sig.c, C.c_int32(f64), C.c_int32(ndet), C.c_int64(nticks), C.c_int32(cpt), C.c_int32(sf),
                                             thr.c, cm.c, C.c_int32(nmod), C.c_int64(digit_ticks), C.c_int32(max_trig),
                                             C.c_void_p(tidx.data_ptr()), C.c_void_p(ntr.data_ptr()), _l.stream()), "light_get_triggers")
    ntr_h, tidx_h = ntr.cpu().numpy(), tidx.cpu().numpy()
    trig, chans, kinds = [], [], []
    for slot in range(nmod):
        if ntr_h[slot] > max_trig:
            raise RuntimeError("get_triggers: more triggers than provisioned")
        for k in range(int(ntr_h[slot])):
            trig.append(int(tidx_h[slot, k])); chans.append(mod_channels[slot]); kinds.append(0)
    if trig:
        return np.array(trig), np.array(chans), np.array(kinds)
    return np.empty((0,), dtype=int), np.empty((0, len(op)), dtype=int), np.empty((0,), dtype=int)


def gen_light_detector_noise(shape, light_det_noise):
    """``gen_light_detector_noise(shape, light_det_noise)`` (light_sim.py:339-377): noise waveforms with the given
    spectrum and uniformly random phases.  The reference draws the phases with ``cupy.random`` (unseeded, unpinned);
    here they come from torch's CUDA generator and the inverse FFT is torch's (cuFFT) -- same distribution, different
    stream.  Returns a torch CUDA tensor [shape[0], shape[1]] (float64)."""
    import torch
    li = _consts.provider().light
    if not shape[0]:
        return torch.empty(tuple(shape), dtype=torch.float64, device="cuda")
    spec = torch.as_tensor(np.asarray(_host_ids(light_det_noise), dtype=np.float64), device="cuda")
    noise_freq = torch.fft.rfftfreq((spec.shape[-1] - 1) * 2, d=li.LIGHT_DET_NOISE_SAMPLE_SPACING, dtype=torch.float64, device="cuda")
    desired = torch.fft.rfftfreq(int(shape[-1]), d=li.LIGHT_TICK_SIZE, dtype=torch.float64, device="cuda")
    bin_size = torch.diff(desired).mean() if desired.numel() > 1 else torch.tensor(1.0, dtype=torch.float64, device="cuda")
    # linear interpolation of each spectrum at the desired frequencies, 0 outside (cp.interp(..., left=0, right=0))
    idx = torch.searchsorted(noise_freq, desired, right=True).clamp(1, noise_freq.numel() - 1)
    f0, f1 = noise_freq[idx - 1], noise_freq[idx]
    w = (desired - f0) / (f1 - f0)
    ns = spec[:, idx - 1] * (1 - w) + spec[:, idx] * w
    ns = torch.where((desired < noise_freq[0]) | (desired > noise_freq[-1]), torch.zeros_like(ns), ns)
    ns = ns * torch.sqrt(torch.diff(noise_freq).mean() / bin_size) * li.LIGHT_DIGIT_SAMPLE_SPACING / li.LIGHT_TICK_SIZE
    phase = torch.rand(ns.shape, dtype=torch.float64, device="cuda")
    z = ns * torch.exp(2j * np.pi * phase)
    q = 2.0 ** (16 - li.LIGHT_NBIT)
    if shape[1] < 2:
        out = torch.round(z.real) * q
    else:
        out = torch.round(torch.fft.irfft(z, dim=-1)) * q
    if out.shape[1] < shape[1]:
        out = torch.cat([out, torch.zeros((out.shape[0], shape[1] - out.shape[1]), dtype=out.dtype, device="cuda")], dim=-1)
    return out[:, :shape[1]]


def _digitize(sig, f64, nticks, row_chan, row_src, front, padded_len, array_f32, true_id, true_ph, M, trig_chan, digit_samples, truncate,
              digit, digit_id, digit_ph, M_out):
    p = _consts.provider()
    li, si = p.light, p.sim
    ntrig, ndm = trig_chan.shape
    rc = _l.dev(np.ascontiguousarray(row_chan, dtype=np.int64), want=np.int64)
    rs = _l.dev(np.ascontiguousarray(row_src, dtype=np.int32), want=np.int32)
    tc = _l.dev(np.ascontiguousarray(trig_chan, dtype=np.int64), want=np.int64)
    _l.check(_l.lib().lsb_light_digitize(sig.c, C.c_int32(f64), C.c_int64(nticks), C.c_int32(len(row_chan)), rc.c, rs.c, C.c_int64(front),
                                         C.c_int64(padded_len), C.c_int32(1 if array_f32 else 0), true_id.c if M else None,
                                         true_ph.c if M else None, C.c_int32(M), C.c_int64(ntrig), tc.c, C.c_int32(ndm),
                                         C.c_int32(digit_samples), C.c_double(li.LIGHT_DIGIT_SAMPLE_SPACING), C.c_double(li.LIGHT_TICK_SIZE),
                                         C.c_double(si.MC_TRUTH_THRESHOLD), C.c_int32(int(li.LIGHT_NBIT)), C.c_int32(1 if truncate else 0),
                                         digit.c, digit_id.c if M_out else None, digit_ph.c if M_out else None, C.c_int32(M_out),
                                         _l.stream()), "light_digitize")


def sim_triggers(bpg, tpb, signal, signal_op_channel_idx, signal_true_track_id, signal_true_photons, trigger_idx, op_channel_idx,
                 digit_samples, light_det_noise):
    """``sim_triggers(bpg, tpb, signal, signal_op_channel_idx, signal_true_track_id, signal_true_photons, trigger_idx,
    op_channel_idx, digit_samples, light_det_noise)`` (light_sim.py:545-619) -> ``(digit_signal f8[ntrigs, ndet_module,
    digit_samples], truth ids, truth photons)`` as torch CUDA tensors.  With an all-zero noise spectrum the result is the
    reference's bit for bit and the padded / re-ordered waveform arrays are never materialised; with noise, the padded
    array is built on the device and the noise of :func:`gen_light_detector_noise` is added first."""
    import torch
    from math import ceil
    li = _consts.provider().light
    trig = np.asarray(_host_ids(trigger_idx)).astype(np.int64).reshape(-1)
    trig_chan = np.asarray(_host_ids(op_channel_idx)).astype(np.int64)
    trig_chan = trig_chan.reshape(len(trig), -1) if trig_chan.ndim != 2 else trig_chan
    chan = np.asarray(_host_ids(signal_op_channel_idx)).astype(np.int64)
    ti, tp, M = _truth(signal_true_track_id, signal_true_photons, False, "signal")
    ntrig, ndm = len(trig), trig_chan.shape[-1]
    id_dtype = torch.int64
    digit = torch.zeros((ntrig, ndm, int(digit_samples)), dtype=torch.float64, device="cuda")
    digit_id = torch.full((ntrig, ndm, int(digit_samples), M), -1, dtype=id_dtype, device="cuda")
    digit_ph = torch.zeros((ntrig, ndm, int(digit_samples), M), dtype=torch.float64, device="cuda")
    if ntrig == 0:
        return digit, digit_id, digit_ph
    sig, f64 = _wave(signal)
    nsig_in, nticks = sig.shape
    pre = int(ceil(li.LIGHT_TRIG_WINDOW[0] / li.LIGHT_TICK_SIZE))
    front = int(pre - trig.min()) if trig.min() - pre < 0 else 0
    post = int(ceil(li.LIGHT_TRIG_WINDOW[1] / li.LIGHT_TICK_SIZE))
    back = max(int(post + (trig + front).max() - (nticks + front)), 0)
    L = front + nticks + back
    missing = np.unique(trig_chan[~np.isin(trig_chan, chan)])
    rows_chan = np.concatenate([chan, missing])
    rows_src = np.concatenate([np.arange(nsig_in), np.full(len(missing), -1)])
    if len(missing):
        order = np.argsort(rows_chan, kind="stable")
        rows_chan, rows_src = rows_chan[order], rows_src[order]
    noise = None if light_det_noise is None else np.asarray(_host_ids(light_det_noise))
    noisy = noise is not None and noise.size and np.any(noise != 0)
    D = lambda t: _l.Dev(t.data_ptr(), tuple(t.shape), np.dtype(str(t.dtype).replace("torch.", "")), keep=t)
    if not noisy:
        array_f32 = (front == 0 and back == 0 and len(missing) == 0 and not f64)
        _digitize(sig, f64, nticks, rows_chan, rows_src, front, L, array_f32, ti, tp, M, trig_chan, int(digit_samples), True,
                  D(digit), D(digit_id), D(digit_ph), M)
        return digit, digit_id, digit_ph
    # noise: materialise the padded, channel-sorted waveforms (float64 as in the reference once anything is concatenated)
    st = torch.as_tensor(signal, device="cuda") if not isinstance(signal, np.ndarray) else torch.from_numpy(signal).cuda()
    full = torch.zeros((len(rows_chan), L), dtype=torch.float64 if (front or back or len(missing) or f64) else torch.float32, device="cuda")
    have = rows_src >= 0
    full[torch.from_numpy(np.nonzero(have)[0]).cuda(), front:front + nticks] = st[torch.from_numpy(rows_src[have]).cuda()].to(full.dtype)
    full += gen_light_detector_noise(full.shape, noise[rows_chan]).to(full.dtype)
    # truth arrays keep their own (unpadded) description: rows_src / front still apply to them
    fd = D(full)
    p = _consts.provider()
    rc = _l.dev(np.ascontiguousarray(rows_chan, dtype=np.int64), want=np.int64)
    # the signal rows are now already in final order: identity source map for the waveform, original map for the truth
    # (handled by materialising the truth arrays the same way when present)
    if M:
        tid_full = torch.full((len(rows_chan), L, M), -1, dtype=torch.int64, device="cuda")
        tph_full = torch.zeros((len(rows_chan), L, M), dtype=torch.float64, device="cuda")
        tt_i = torch.as_tensor(signal_true_track_id, device="cuda") if not isinstance(signal_true_track_id, np.ndarray) else torch.from_numpy(signal_true_track_id).cuda()
        tt_p = torch.as_tensor(signal_true_photons, device="cuda") if not isinstance(signal_true_photons, np.ndarray) else torch.from_numpy(signal_true_photons).cuda()
        dst = torch.from_numpy(np.nonzero(have)[0]).cuda(); src = torch.from_numpy(rows_src[have]).cuda()
        tid_full[dst, front:front + nticks] = tt_i[src].to(torch.int64)
        tph_full[dst, front:front + nticks] = tt_p[src].to(torch.float64)
        ti, tp = D(tid_full), D(tph_full)
    _digitize(fd, 1 if full.dtype == torch.float64 else 0, L, rows_chan, np.arange(len(rows_chan)), 0, L, full.dtype == torch.float32,
              ti, tp, M, trig_chan, int(digit_samples), True, D(digit), D(digit_id), D(digit_ph), M)
    return digit, digit_id, digit_ph


@_l.kernel
def digitize_signal(signal, signal_op_channel_idx, trigger_idx, trigger_op_channel_idx, signal_true_track_id, signal_true_photons,
                    digit_signal, digit_signal_true_track_id, digit_signal_true_photons):
    """``digitize_signal[BPG, TPB](signal, signal_op_channel_idx, trigger_idx, trigger_op_channel_idx, signal_true_track_id,
    signal_true_photons, digit_signal, digit_signal_true_track_id, digit_signal_true_photons)`` (light_sim.py:480-543): the
    raw kernel -- no padding, no rounding; outputs are the caller's pre-filled arrays."""
    sig, f64 = _wave(signal)
    chan = np.asarray(_host_ids(signal_op_channel_idx)).astype(np.int64)
    trig_chan = np.asarray(_host_ids(trigger_op_channel_idx)).astype(np.int64)
    ti, tp, M = _truth(signal_true_track_id, signal_true_photons, False, "signal")
    dg = _l.dev(digit_signal, want=np.float64, write=True, name="digit_signal")
    di, dp, _ = _truth(digit_signal_true_track_id, digit_signal_true_photons, True, "digit_signal")
    M_out = di.shape[-1] if len(di.shape) == 4 else 0
    ntrig, ndm, ns = dg.shape
    nsig, nticks = sig.shape
    _digitize(sig, f64, nticks, chan, np.arange(nsig), 0, nticks, not f64, ti, tp, M, trig_chan.reshape(ntrig, ndm), ns, False,
              dg, di, dp, M_out)
    _l.finish(dg, di, dp)


# ---------------------------------------------------------------------------------------------------------
# zero suppression of the waveform truth (the compaction in front of the light HDF5 export)
# ---------------------------------------------------------------------------------------------------------
TRUTH_DTYPE = np.dtype([('trigger_id', 'i4'), ('op_channel_id', 'i4'), ('tick', 'i4'), ('event_id', 'i4'), ('segment_id', 'i8'),
                        ('pe_current', 'f8')])


def zero_suppress_waveform_truth(waveforms_true_track_id, waveforms_true_photons, i_evt, i_trig, i_mod=-1):
    """``zero_suppress_waveform_truth(ids, photons, i_evt, i_trig, i_mod=-1)`` (light_sim.py:621-661): the truth slots that are
    not -1, flattened to records ``['trigger_id', 'op_channel_id', 'tick', 'event_id', 'segment_id', 'pe_current']`` in the
    reference's enumeration order (a Python loop over every slot there; flag -> prefix sum -> scatter here).  NumPy out."""
    import torch
    light = _consts.provider().light
    chan = np.asarray(light.TPC_TO_OP_CHANNEL)
    chan = (chan[(i_mod - 1) * 2:i_mod * 2] if i_mod > 0 else chan[:]).ravel()
    ids = _l.dev(waveforms_true_track_id, want=np.int64, name="waveforms_true_track_id")
    ph = _l.dev(waveforms_true_photons, want=np.float64, name="waveforms_true_photons")
    if len(ids.shape) != 4 or tuple(ph.shape) != tuple(ids.shape):
        raise ValueError("truth arrays must have shape (ntrigs, ndet, nsamples, ntruth)")
    nt, nd, ns, M = (int(x) for x in ids.shape)
    if nd > chan.size:
        raise IndexError("index %d is out of bounds for axis 0 with size %d" % (nd - 1, chan.size))
    n = nt * nd * ns * M
    lib = _l.lib()
    lib.lsb_light_truth_ws_bytes.restype = C.c_int64
    nws = int(lib.lsb_light_truth_ws_bytes(C.c_int64(n)))
    ws = torch.empty(max(nws, 16), dtype=torch.uint8, device="cuda")
    rows = torch.empty((max(n, 1), TRUTH_DTYPE.itemsize), dtype=torch.uint8, device="cuda")
    cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
    cd = torch.from_numpy(np.ascontiguousarray(chan[:max(nd, 1)], dtype=np.int32)).cuda()
    _l.check(lib.lsb_light_zero_suppress_truth(ids.c, ph.c, C.c_int64(nt), C.c_int32(nd), C.c_int32(ns), C.c_int32(M),
                                               C.c_void_p(cd.data_ptr()), C.c_int32(int(i_evt)), C.c_int32(int(i_trig)),
                                               C.c_void_p(rows.data_ptr()), C.c_void_p(cnt.data_ptr()), C.c_void_p(ws.data_ptr()),
                                               C.c_int64(nws), _l.stream()), "light_zero_suppress_truth")
    k = int(cnt.item())
    return rows[:k].cpu().numpy().reshape(-1).view(TRUTH_DTYPE).copy() if k else np.empty(0, dtype=TRUTH_DTYPE)
