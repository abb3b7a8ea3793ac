"""CPU restatement of the light trigger + digitisation stage -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.  It restates, with explicit
loops, what ``larndsim/light_sim.py`` computes in

* ``get_triggers`` (:380-477): per trigger group the float32 channel sum, its average over blocks of
  ``round(LIGHT_DIGIT_SAMPLE_SPACING / LIGHT_TICK_SIZE)`` ticks (float64, zero padded), the comparison with the
  group threshold, the per-module OR and the sequential trigger search with its dead time -- including the
  reference's index bookkeeping (the remaining waveform is re-sliced by an absolute index, :441-452), which makes
  later triggers of a module drift; it is replicated, not corrected;
* ``sim_triggers`` (:545-619) with a zero noise spectrum (the noise itself is ``cupy.random``: unpinned) and
  ``digitize_signal`` (:480-543) with ``interp`` (:241-271): front / back zero padding, channels without a
  simulated waveform, linear interpolation at ``isample * spacing / tick`` (the trigger tick does not enter: the
  reference's offset is commented out, :494), truth propagation and the final rounding to ``LIGHT_NBIT`` bits.

Pinned: tools/gen_golden_light_trigger.py runs the reference's own functions (NumPy standing in for CuPy, Numba's
CUDA simulator for the kernel) and commits inputs and outputs under tests/golden/light_trigger_*.npz.
"""
from math import ceil, floor

import numpy as np


def block_means(signal, cpt, sample_factor):
    """[ngrp, nblk] float64: mean over `sample_factor` ticks (zero padded at the end) of the float32 sum of the
    `cpt` channels of every group.  NumPy reductions are used so the rounding order is NumPy's."""
    ndet, nticks = signal.shape
    ngrp = ndet // cpt
    gsum = np.zeros((ngrp, nticks), dtype=signal.dtype)
    for g in range(ngrp):
        acc = signal[g * cpt].copy()
        for c in range(1, cpt):
            acc = acc + signal[g * cpt + c]                      # sequential, in the signal's own precision
        gsum[g] = acc
    padding = sample_factor - nticks % sample_factor             # 1 .. sample_factor (a full block when it divides)
    nblk = (nticks + padding) // sample_factor
    padded = np.zeros((ngrp, nblk * sample_factor), dtype=np.float64)
    padded[:, :nticks] = gsum
    return np.add.reduce(padded.reshape(ngrp, nblk, sample_factor), axis=-1) / sample_factor


def get_triggers(signal, group_threshold, op_channel_idx, i_subbatch, C):
    """``C``: dict with OP_CHANNEL_PER_TRIG, LIGHT_DIGIT_SAMPLE_SPACING, LIGHT_TICK_SIZE, LIGHT_TRIG_WINDOW, LIGHT_TRIG_MODE,
    OP_CHANNEL_TO_TPC, TPC_TO_OP_CHANNEL, TPC_TO_MODULE {tpc: module}, MODULE_TO_TPCS {module: [tpc, ...]}."""
    ndet, nticks = signal.shape
    cpt = C["OP_CHANNEL_PER_TRIG"]
    sf = round(C["LIGHT_DIGIT_SAMPLE_SPACING"] / C["LIGHT_TICK_SIZE"])
    means = block_means(signal, cpt, sf)
    digit_ticks = ceil((C["LIGHT_TRIG_WINDOW"][1] + C["LIGHT_TRIG_WINDOW"][0]) / C["LIGHT_TICK_SIZE"])
    op_channel_idx = np.asarray(op_channel_idx)
    trig, chans, kinds = [], [], []
    if C["LIGHT_TRIG_MODE"] == 0:
        tpcs = np.unique(np.asarray(C["OP_CHANNEL_TO_TPC"])[op_channel_idx])
        mods = np.unique([C["TPC_TO_MODULE"][int(t)] for t in tpcs])
        for mod in mods:
            mod_channels = np.asarray(C["TPC_TO_OP_CHANNEL"])[C["MODULE_TO_TPCS"][int(mod)]].ravel()
            in_module = np.isin(op_channel_idx, mod_channels)
            above = np.zeros(nticks, dtype=bool)
            for d in np.nonzero(in_module)[0]:
                g = d // cpt
                above |= np.repeat(means[g] < group_threshold[g], sf)[:nticks]
            base = 0                 # absolute tick of the first element of the remaining waveform
            last_trigger = 0         # the reference's running offset
            while base < nticks and above[base:].any():
                rel = int(np.argmax(above[base:]))
                idx = rel + last_trigger
                trig.append(idx); chans.append(mod_channels); kinds.append(0)
                base += idx + digit_ticks            # the remaining waveform is cut at the absolute index
                last_trigger = idx + digit_ticks
    elif C["LIGHT_TRIG_MODE"] == 1 and i_subbatch == 0:
        trig.append(0); chans.append(op_channel_idx); kinds.append(1)
    if trig:
        return np.array(trig), np.array(chans), np.array(kinds)
    return np.empty((0,), dtype=int), np.empty((0, len(op_channel_idx)), dtype=int), np.empty((0,), dtype=int)


def interp(idx, arr, low, high):
    i0 = int(floor(idx))
    if i0 < 0:
        return low
    if i0 > len(arr) - 1:
        return high
    if i0 == idx:
        return float(arr[i0])
    if i0 > len(arr) - 2:
        return high
    v0, v1 = arr[i0], arr[i0 + 1]
    # compiled (Numba) typing: the difference is taken in the array's precision, the product and sum in float64
    d = np.subtract(v1, v0) if isinstance(v0, np.floating) else v1 - v0
    return float(v0) + float(d) * (idx - i0)


def sim_triggers(signal, signal_op_channel_idx, true_track_id, true_photons, trigger_idx, op_channel_idx, digit_samples, C):
    """Zero-noise ``sim_triggers``: (digit_signal f8[ntrig, ndet_module, nsamples], truth ids, truth photons).
    ``C`` additionally holds LIGHT_NBIT and MC_TRUTH_THRESHOLD."""
    ntrig, ndm = trigger_idx.shape[0], op_channel_idx.shape[-1]
    M = true_track_id.shape[-1]
    out = np.zeros((ntrig, ndm, digit_samples), dtype=np.float64)
    out_id = np.full((ntrig, ndm, digit_samples, M), -1, dtype=true_track_id.dtype)
    out_ph = np.zeros((ntrig, ndm, digit_samples, M), dtype=true_photons.dtype)
    if ntrig == 0:
        return out, out_id, out_ph
    tick = C["LIGHT_TICK_SIZE"]
    signal = np.asarray(signal)
    chan = np.asarray(signal_op_channel_idx).copy()
    # zero padding in front / behind (the trigger index itself is not used by the interpolation)
    pre = int(ceil(C["LIGHT_TRIG_WINDOW"][0] / tick))
    padded_trig = trigger_idx.copy()
    front = int(pre - trigger_idx.min()) if trigger_idx.min() - pre < 0 else 0
    padded_trig = padded_trig + front
    post = int(ceil(C["LIGHT_TRIG_WINDOW"][1] / tick))
    back = int(post + padded_trig.max() - (signal.shape[1] + front))
    back = back if back > 0 else 0
    promote = front > 0 or back > 0
    sig = np.zeros((signal.shape[0], front + signal.shape[1] + back), dtype=np.float64 if promote else signal.dtype)
    sig[:, front:front + signal.shape[1]] = signal
    tid = np.full((signal.shape[0], sig.shape[1], M), -1, dtype=true_track_id.dtype)
    tph = np.zeros((signal.shape[0], sig.shape[1], M), dtype=true_photons.dtype)
    tid[:, front:front + signal.shape[1]] = true_track_id
    tph[:, front:front + signal.shape[1]] = true_photons
    # channels that are read out but were not simulated: zero rows, then everything ordered by channel id
    missing = np.unique(op_channel_idx[~np.isin(op_channel_idx, chan)])
    if len(missing):
        sig = np.concatenate([sig.astype(np.float64), np.zeros((len(missing), sig.shape[1]))], axis=0)
        tid = np.concatenate([tid, np.full((len(missing),) + tid.shape[1:], -1, dtype=tid.dtype)], axis=0)
        tph = np.concatenate([tph, np.zeros((len(missing),) + tph.shape[1:], dtype=tph.dtype)], axis=0)
        chan = np.concatenate([chan, missing])
        order = np.argsort(chan, kind="stable")
        sig, tid, tph, chan = sig[order], tid[order], tph[order], chan[order]
    nsig = sig.shape[0]
    thr = C["MC_TRUTH_THRESHOLD"]
    for it in range(ntrig):
        for im in range(ndm):
            idet = op_channel_idx[it, im]
            isig = nsig - 1                                      # the reference's search leaves the last row when nothing matches
            for k in range(nsig):
                if chan[k] == idet:
                    isig = k
                    break
            for s in range(digit_samples):
                st = s * C["LIGHT_DIGIT_SAMPLE_SPACING"] / tick
                out[it, im, s] = interp(st, sig[isig], 0, 0)
                t0, t1 = int(floor(st)), int(ceil(st))
                n = 0
                for j in range(M):
                    if n >= M:
                        break
                    if tid[isig, t0, j] == -1:
                        break
                    p0, p1 = 0, 0
                    if tid[isig, t0, j] == out_id[it, im, s, n] or out_id[it, im, s, n] == -1:
                        out_id[it, im, s, n] = tid[isig, t0, j]
                        n += 1
                        p0 = tph[idet, t0, j]                   # (sic) indexed with the channel id
                        if abs(p0) < thr:
                            continue
                        if tid[isig, t0, j] == tid[isig, t1, j]:
                            p1 = tph[isig, t1, j]
                        else:
                            for k in range(M):
                                if tid[isig, t0, j] == tid[isig, t1, k]:
                                    p1 = tph[isig, t1, k]
                                    break
                    if out_id[it, im, s, n - 1] != -1:
                        out_ph[it, im, s, n - 1] = interp(st - t0, (p0, p1), 0, 0)
    q = 2 ** (16 - C["LIGHT_NBIT"])
    return np.round(out / q) * q, out_id, out_ph


# ---------------------------------------------------------------------------------------------------------
# extent of the light window: get_nticks (larndsim/light_sim.py:24-42), get_active_op_channel (:44-57)
# pinned by tools/gen_golden_light_extent.py -> tests/golden/light_extent.npz
# ---------------------------------------------------------------------------------------------------------
def lit_extremes(light_incidence):
    """(earliest, latest first-photon time among entries with photons, per-channel 'lit' flags), explicit loops"""
    S, ndet = light_incidence.shape
    lo = hi = None
    lit = np.zeros(ndet, dtype=bool)
    nph, t0 = light_incidence["n_photons_det"], light_incidence["t0_det"]
    for s in range(S):
        for d in range(ndet):
            if nph[s, d] > 0:
                lit[d] = True
                v = t0[s, d]
                lo = v if lo is None or v < lo else lo
                hi = v if hi is None or v > hi else hi
    return lo, hi, lit


def get_nticks(light_incidence, C):
    """``C``: LIGHT_TRIG_MODE, LIGHT_WINDOW, LIGHT_TICK_SIZE.  float32 extremes combined with Python floats stay float32."""
    lo, hi, lit = lit_extremes(light_incidence)
    if lit.any() and C["LIGHT_TRIG_MODE"] == 0:
        start = lo - C["LIGHT_WINDOW"][0]
        stop = hi + C["LIGHT_WINDOW"][1]
        return int(np.ceil((stop - start) / C["LIGHT_TICK_SIZE"])), start
    return int((C["LIGHT_WINDOW"][1] + C["LIGHT_WINDOW"][0]) / C["LIGHT_TICK_SIZE"]), 0


def get_active_op_channel(light_incidence):
    return np.nonzero(lit_extremes(light_incidence)[2])[0].astype(np.int32)


# ---------------------------------------------------------------------------------------------------------
# zero_suppress_waveform_truth (larndsim/light_sim.py:621-661); pinned by tools/gen_golden_light_extent.py
# ---------------------------------------------------------------------------------------------------------
TRUTH_DTYPE = np.dtype([("trigger_id", "i4"), ("op_channel_id", "i4"), ("tick", "i4"), ("event_id", "i4"), ("segment_id", "i8"),
                        ("pe_current", "f8")])


def zero_suppress_waveform_truth(true_track_id, true_photons, i_evt, i_trig, op_channel):
    """``op_channel``: channel id of every column (TPC_TO_OP_CHANNEL of the module, flattened).  The trigger id is a running
    sum: every kept slot adds its trigger index to it before it is recorded (:644)."""
    nt, nd, ns, M = true_track_id.shape
    rows = []
    running = i_trig
    for t in range(nt):
        for d in range(nd):
            for s in range(ns):
                for m in range(M):
                    if true_track_id[t, d, s, m] != -1:
                        running += t
                        rows.append((running, op_channel[d], s, i_evt, true_track_id[t, d, s, m], true_photons[t, d, s, m]))
    out = np.empty(len(rows), dtype=TRUTH_DTYPE)
    for k, r in enumerate(rows):
        out[k] = r
    return out
