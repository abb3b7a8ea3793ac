"""Test infrastructure: ctypes wrapper of the CPU oracle (oracle/liblarnd_oracle.so) and the
comparison of the CUDA chain with it.  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
legs import this; nothing here is on the product path."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from larndsim_b200 import _abi, consts as lconsts, synth  # noqa: E402

ORACLE_LIB = os.path.join(ROOT, "oracle", "liblarnd_oracle.so")
_orc = None


def oracle_lib():
    global _orc
    if _orc is None:
        if not os.path.exists(ORACLE_LIB):
            subprocess.check_call(["make", "-s", "-C", os.path.join(ROOT, "oracle")])
        _orc = C.CDLL(ORACLE_LIB)
        _orc.orc_unique_pixels.restype = C.c_int64
    return _orc


def P(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def rng_states(n, seed, subsequence_start=0):
    st = np.zeros((int(n), 2), dtype=np.uint64)
    oracle_lib().orc_rng_create_states(P(st), C.c_int64(int(n)), C.c_uint64(int(seed)), C.c_uint64(int(subsequence_start)))
    return st


class Oracle:
    """NumPy-in / NumPy-out calls of the C restatement; constants from an ``lsb_consts`` snapshot."""

    def __init__(self, c=None):
        self.c = c if c is not None else lconsts.snapshot()
        self.lib = oracle_lib()

    def _L(self, tracks):
        return _abi.track_layout(tracks.dtype)

    def quench(self, tracks, mode):
        L = self._L(tracks)
        return self.lib.orc_quench(C.byref(self.c), C.byref(L), P(tracks), C.c_int64(len(tracks)), C.c_int32(int(mode)))

    def drift(self, tracks):
        L = self._L(tracks)
        return self.lib.orc_drift(C.byref(self.c), C.byref(L), P(tracks), C.c_int64(len(tracks)))

    def max_pixels(self, tracks):
        L = self._L(tracks)
        out = np.zeros(1, dtype=np.int64)
        self.lib.orc_max_pixels(C.byref(self.c), C.byref(L), P(tracks), C.c_int64(len(tracks)), P(out))
        return int(out[0])

    def get_pixels(self, tracks, max_active, P_, radius):
        L = self._L(tracks)
        S = len(tracks)
        act = np.full((S, max_active), -1, dtype=np.int32)
        nb = np.full((S, P_), -1, dtype=np.int32)
        nr = np.full((S, P_), -1, dtype=np.int32)
        npl = np.zeros(S, dtype=np.float64)
        self.lib.orc_get_pixels(C.byref(self.c), C.byref(L), P(tracks), C.c_int64(S), P(act), C.c_int32(max_active), P(nb), P(nr),
                                C.c_int32(P_), P(npl), C.c_int32(radius))
        return act, nb, nr, npl

    def time_intervals(self, tracks):
        L = self._L(tracks)
        ts = np.zeros(len(tracks), dtype=np.float64)
        tm = np.zeros(1, dtype=np.int64)
        self.lib.orc_time_intervals(C.byref(self.c), C.byref(L), P(tracks), C.c_int64(len(tracks)), P(ts), P(tm))
        return ts, int(tm[0])

    def unique_pixels(self, pixels):
        flat = np.ascontiguousarray(pixels.reshape(-1))
        out = np.zeros(flat.size, dtype=np.int32)
        n = self.lib.orc_unique_pixels(P(flat), C.c_int64(flat.size), P(out))
        return out[:n].copy()

    def pixel_index_map(self, pixels, uniq):
        flat = np.ascontiguousarray(pixels.reshape(-1))
        out = np.zeros(flat.size, dtype=np.int64)
        self.lib.orc_pixel_index_map(P(flat), C.c_int64(flat.size), P(uniq), C.c_int64(uniq.size), P(out))
        return out.reshape(pixels.shape)

    def tracks_current_mc(self, tracks, pixels, T, response, states, mode):
        L = self._L(tracks)
        S, P_ = pixels.shape
        sig = np.zeros((S, P_, T), dtype=np.float32)
        r = np.ascontiguousarray(response)
        self.lib.orc_tracks_current_mc(C.byref(self.c), C.byref(L), P(tracks), C.c_int64(S), P(pixels), C.c_int32(P_), P(sig),
                                       C.c_int32(T), P(r), C.c_int32(r.shape[0]), C.c_int32(r.shape[1]), C.c_int32(r.shape[2]),
                                       C.c_int32(1 if r.dtype == np.float64 else 0), P(states), C.c_int32(mode))
        return sig

    def tracks_current(self, tracks, pixels, T, response):
        L = self._L(tracks)
        S, P_ = pixels.shape
        sig = np.zeros((S, P_, T), dtype=np.float32)
        r = np.ascontiguousarray(response)
        self.lib.orc_tracks_current(C.byref(self.c), C.byref(L), P(tracks), C.c_int64(S), P(pixels), C.c_int32(P_), P(sig),
                                    C.c_int32(T), P(r), C.c_int32(r.shape[0]), C.c_int32(r.shape[1]), C.c_int32(r.shape[2]),
                                    C.c_int32(1 if r.dtype == np.float64 else 0))
        return sig

    def track_pixel_map(self, uniq, pixels, K):
        tpm = np.full((uniq.size, K), -1, dtype=np.int64)
        self.lib.orc_get_track_pixel_map(P(tpm), C.c_int32(K), P(uniq), C.c_int64(uniq.size), P(pixels),
                                         C.c_int64(pixels.shape[0]), C.c_int32(pixels.shape[1]))
        return tpm

    def track_pixel_map2(self, uniq, pixels, dist, max_distance, K):
        tpm = np.full((uniq.size, K), -1, dtype=np.int64)
        self.lib.orc_get_track_pixel_map2(P(tpm), C.c_int32(K), P(uniq), C.c_int64(uniq.size), P(pixels), P(dist),
                                          C.c_int64(pixels.shape[0]), C.c_int32(pixels.shape[1]), C.c_int32(max_distance))
        return tpm

    def sum_pixel_signals(self, signals, track_starts, pim, tpm, Tt):
        U, K = tpm.shape
        S, P_, T = signals.shape
        ps = np.zeros((U, Tt), dtype=np.float64)
        pts = np.zeros((U, Tt, K), dtype=np.float64)
        of = np.zeros(U, dtype=np.float64)
        self.lib.orc_sum_pixel_signals(C.byref(self.c), P(ps), C.c_int64(U), C.c_int32(Tt), P(signals), C.c_int64(S), C.c_int32(P_),
                                       C.c_int32(T), P(track_starts), P(pim), P(tpm), C.c_int32(K), P(pts), P(of))
        return ps, pts, of

    def get_adc_values(self, ps, pts, time_ticks, A, time_padding, states, thresholds):
        U, Tt = ps.shape
        K = pts.shape[2]
        adc = np.zeros((U, A), dtype=np.float64)
        ticks = np.zeros((U, A), dtype=np.float64)
        cf = np.zeros((U, A, K), dtype=np.float64)
        self.lib.orc_get_adc_values(C.byref(self.c), P(ps), P(pts), C.c_int64(U), C.c_int32(Tt), C.c_int32(K), P(time_ticks),
                                    C.c_int32(time_ticks.size), P(adc), P(ticks), C.c_int32(A), C.c_double(time_padding),
                                    P(states), P(cf), P(thresholds))
        return adc, ticks, cf

    def digitize(self, q, gain=None):
        q = np.ascontiguousarray(q, dtype=np.float64)
        out = np.zeros_like(q)
        g = None if gain is None else np.ascontiguousarray(np.broadcast_to(gain, q.shape), dtype=np.float64)
        self.lib.orc_digitize(C.byref(self.c), P(q), P(g), C.c_int64(q.size), P(out))
        return out


class OracleLight:
    """Light-chain calls of the oracle (lightLUT.py / light_sim.py restatements)."""

    def __init__(self, c=None):
        self.c = c if c is not None else lconsts.snapshot()
        self.lib = oracle_lib()

    def light_incidence(self, tracks, lut, ndet, eff, ch2tpc):
        L = _abi.track_layout(tracks.dtype)
        LL = _abi.lut_layout(lut.dtype, lut.shape)
        linc = np.zeros((len(tracks), ndet), dtype=[("segment_id", "u4"), ("n_photons_det", "f4"), ("t0_det", "f4")])
        LI = _abi.linc_layout(linc.dtype)
        vox = np.zeros((len(tracks), 3), dtype=np.int32)
        eff = np.ascontiguousarray(eff, dtype=np.float64)
        ch2tpc = np.ascontiguousarray(ch2tpc, dtype=np.int64)
        self.lib.orc_calculate_light_incidence(C.byref(self.c), C.byref(L), P(tracks), C.c_int64(len(tracks)), P(lut), C.byref(LL),
                                               P(linc), C.byref(LI), C.c_int32(ndet), P(vox), P(eff), P(ch2tpc))
        return linc, vox

    def sum_light_signals(self, tracks, vox, seg_ids, linc, op_channel, lut, t_start, nticks, n_true, sorted_idx, prof_len):
        L = _abi.track_layout(tracks.dtype)
        LL = _abi.lut_layout(lut.dtype, lut.shape)
        LI = _abi.linc_layout(linc.dtype)
        nd = len(op_channel)
        inc = np.zeros((nd, nticks), dtype=np.float32)
        tid = np.full((nd, nticks, n_true), -1, dtype=np.int64)
        tph = np.zeros((nd, nticks, n_true), dtype=np.float64)
        self.lib.orc_sum_light_signals(C.byref(self.c), C.byref(L), P(tracks), C.c_int64(len(tracks)), P(vox), P(seg_ids), P(linc),
                                       C.byref(LI), C.c_int32(linc.shape[1]), P(op_channel), P(lut), C.byref(LL),
                                       C.c_double(float(t_start)), P(inc), C.c_int32(nd), C.c_int32(nticks), P(tid), P(tph),
                                       C.c_int32(n_true), P(sorted_idx), C.c_int64(sorted_idx.shape[1]), C.c_double(float(prof_len)))
        return inc, tid, tph

    def scintillation(self, inc, tid, tph):
        nd, nticks = inc.shape
        out = np.zeros_like(inc)
        oid = np.full_like(tid, -1)
        oph = np.zeros_like(tph)
        self.lib.orc_calc_scintillation_effect(C.byref(self.c), P(inc), P(tid), P(tph), P(out), P(oid), P(oph), C.c_int32(nd),
                                               C.c_int32(nticks), C.c_int32(tid.shape[2]), C.c_int32(oid.shape[2]))
        return out, oid, oph

    def stat_fluctuations(self, inc, states):
        nd, nticks = inc.shape
        out = np.zeros_like(inc)
        self.lib.orc_calc_stat_fluctuations(C.byref(self.c), P(inc), P(out), C.c_int32(nd), C.c_int32(nticks), P(states))
        return out

    def detector_response(self, inc, tid, tph, gain, impulse):
        nd, nticks = inc.shape
        out = np.zeros_like(inc)
        oid = np.full_like(tid, -1)
        oph = np.zeros_like(tph)
        gain = np.ascontiguousarray(gain, dtype=np.float64)
        impulse = np.ascontiguousarray(impulse, dtype=np.float64)
        self.lib.orc_calc_light_detector_response(C.byref(self.c), P(inc), P(tid), P(tph), P(out), P(oid), P(oph), C.c_int32(nd),
                                                  C.c_int32(nticks), C.c_int32(tid.shape[2]), C.c_int32(oid.shape[2]), P(gain),
                                                  P(impulse), C.c_int32(impulse.size))
        return out, oid, oph


def production_tracks(n, config="module0", seed=12345, kind="cosmic", dtype=None):
    mod = lconsts.load_snapshot(config)
    dt = dtype if dtype is not None else synth.segment_dtype
    if kind == "cosmic":
        return synth.cosmic_segments(n, mod.detector, seed=seed, dtype=dt)
    return synth.beam_spill_segments(n, mod.detector, seed=seed, dtype=dt)


def oracle_front(tracks, orc, quench_mode=2):
    """quench .. time_intervals on the CPU, mirroring cli/simulate_pixels.py:732-1002."""
    c = orc.c
    orc.quench(tracks, quench_mode)
    orc.drift(tracks)
    max_radius = int(np.ceil(max(tracks["tran_diff"]) * 5 / c.pixel_pitch))
    maxpix = orc.max_pixels(tracks)
    P_ = (2 * max_radius + 1) * maxpix + (1 + 2 * max_radius) * max_radius * 2
    act, nb, nr, npl = orc.get_pixels(tracks, maxpix, P_, max_radius)
    uniq = orc.unique_pixels(nb)
    ts, T = orc.time_intervals(tracks)
    return dict(radius=max_radius, maxpix=maxpix, P=P_, active=act, neigh=nb, nrad=nr, npl=npl, uniq=uniq, starts=ts, T=T)


def oracle_back(orc, front, signals, states, n_events=1):
    """pixel_index_map .. digitize on the CPU from given `signals` (cli/simulate_pixels.py:1021-1102)."""
    c = orc.c
    K, A, Tt = c.max_tracks_per_pixel, c.max_adc_values, c.n_time_ticks
    pim = orc.pixel_index_map(front["neigh"], front["uniq"])
    tpm = orc.track_pixel_map2(front["uniq"], front["neigh"], front["nrad"], int(front["nrad"].max()) + 1, K)
    ps, pts, of = orc.sum_pixel_signals(signals, front["starts"], pim, tpm, Tt)
    time_ticks = np.linspace(0, n_events * c.time_interval[1], Tt + 1)
    thr = np.full(len(front["uniq"]), c.discrimination_threshold * c.unit_e)
    adc, ticks, cf = orc.get_adc_values(ps, pts, time_ticks, A, 0.0, states, thr)
    return dict(pim=pim, tpm=tpm, ps=ps, pts=pts, overflow=of, adc=adc, ticks=ticks, cf=cf, digit=orc.digitize(adc))


def oracle_back_chunked(orc, front, signals, states, n_events=1, chunk=2048):
    """`oracle_back` with bounded memory: the reference's dense pixels_tracks_signals is [U, Tt, K] float64 (0.8 MB per pixel for
    module0, 1.3 MB for ND-LAr), so the pixels are processed `chunk` at a time -- pixels are independent in sum_pixel_signals and
    get_adc_values (RNG state index = pixel, fee.py:557), the results are identical to one dense call."""
    c = orc.c
    K, A, Tt = c.max_tracks_per_pixel, c.max_adc_values, c.n_time_ticks
    uniq = front["uniq"]
    U = len(uniq)
    pim = orc.pixel_index_map(front["neigh"], uniq)
    tpm = orc.track_pixel_map2(uniq, front["neigh"], front["nrad"], int(front["nrad"].max()) + 1, K)
    time_ticks = np.linspace(0, n_events * c.time_interval[1], Tt + 1)
    ps = np.zeros((U, Tt)); of = np.zeros(U)
    adc = np.zeros((U, A)); ticks = np.zeros((U, A)); cf = np.zeros((U, A, K))
    for u0 in range(0, U, chunk):
        u1 = min(U, u0 + chunk)
        sel = (pim >= u0) & (pim < u1)
        pim_c = np.where(sel, pim - u0, -1)
        ps_c, pts_c, of_c = orc.sum_pixel_signals(signals, front["starts"], pim_c, np.ascontiguousarray(tpm[u0:u1]), Tt)
        thr = np.full(u1 - u0, c.discrimination_threshold * c.unit_e)
        st_c = np.ascontiguousarray(states[u0:u1])
        adc_c, ticks_c, cf_c = orc.get_adc_values(ps_c, pts_c, time_ticks, A, 0.0, st_c, thr)
        states[u0:u1] = st_c
        ps[u0:u1], of[u0:u1], adc[u0:u1], ticks[u0:u1], cf[u0:u1] = ps_c, of_c, adc_c, ticks_c, cf_c
        del pts_c
    return dict(pim=pim, tpm=tpm, ps=ps, overflow=of, adc=adc, ticks=ticks, cf=cf, digit=orc.digitize(adc))


def rel_err_rows(a, b, rows=256):
    """`rel_err` over the leading axis in slabs (bounded temporaries for GB-sized waveform arrays)"""
    worst = 0.0
    for i in range(0, a.shape[0], rows):
        worst = max(worst, rel_err(a[i:i + rows], b[i:i + rows]))
    return worst


#: The oracle evaluates exp/log/erf with glibc, the CUDA kernels with libdevice (what the reference's
#: Numba-CUDA build calls): both are <= 1 ulp but not bit-identical, exactly like the reference's own GPU
#: and CUDA-simulator builds.  float64 record fields that go through a transcendental are therefore
#: compared to a few ulp; everything stored as float32 / integer must be identical.
F64_RTOL = 1e-14


def records_equal(a, b, f64_rtol=0.0):
    """Field-by-field equality of two structured arrays (padding bytes are not data)."""
    if a.dtype != b.dtype or a.shape != b.shape:
        return False
    for name in a.dtype.names:
        x, y = a[name], b[name]
        if x.dtype.kind == "f":
            if np.array_equal(x, y, equal_nan=True):
                continue
            if x.dtype.itemsize == 8 and f64_rtol > 0 and np.allclose(x, y, rtol=f64_rtol, atol=0, equal_nan=True):
                continue
            return False
        elif not np.array_equal(x, y):
            return False
    return True


def rel_err_peak(a, b):
    """max |a-b| / (|b| + max|b| per waveform) -- allclose(rtol, atol = rtol * peak)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.abs(b) + np.abs(b).max(axis=-1, keepdims=True)
    den[den == 0] = 1.0
    return float((np.abs(a - b) / den).max())


def rel_err(a, b):
    """max |a-b| / (|b| + 1e-2 * max|b| per waveform): elementwise relative error with a floor that keeps
    zero crossings of bipolar waveforms from dominating."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    if a.size == 0:
        return 0.0
    scale = np.abs(b).max(axis=-1, keepdims=True)
    den = np.abs(b) + 1e-2 * scale
    den[den == 0] = 1.0
    return float((np.abs(a - b) / den).max())


def chain_vs_oracle(n_segments=64, config="module0", seed=7, noise=True, kind="cosmic", rng_seed=1, response=None, dense=False,
                    exact_fractions=False):
    """Run the fused CUDA chain on a synthetic batch and compare with the oracle stage by stage."""
    import torch
    from larndsim_b200 import chain as lchain, _launch as ll
    tracks = production_tracks(n_segments, config, seed, kind)
    mod = lconsts.load_snapshot(config)
    if not noise:
        mod.detector.RESET_NOISE_CHARGE = 0
        mod.detector.UNCORRELATED_NOISE_CHARGE = 0
        mod.detector.DISCRIMINATOR_NOISE = 0
    if response is None:
        response = synth.response_lut(mod.detector)
    c = lconsts.snapshot()
    launches0 = ll.lib().lsb_launch_count()
    ch = lchain.Chain(tracks.dtype, response, rng_mode="cloud", dense=dense, exact_fractions=exact_fractions)
    dtr = ll.DeviceRecords(host=tracks)
    res = ch.run(dtr, rng_seed=rng_seed, n_events=1)
    torch.cuda.synchronize()
    launches = ll.lib().lsb_launch_count() - launches0
    g_tracks = dtr.copy_to_host()
    # ---- oracle ----
    orc = Oracle(c)
    otr = tracks.copy()
    front = oracle_front(otr, orc, quench_mode=c.mode_birks)
    S, P_ = front["neigh"].shape
    out = dict(S=S, U=len(front["uniq"]), T=front["T"], launches=int(launches), n_hits=res.n_hits)
    out["tracks_equal"] = records_equal(g_tracks, otr)
    out["shape_equal"] = (res.max_neighbors == P_ and res.n_ticks == front["T"] and res.n_unique_pixels == len(front["uniq"]))
    out["unique_equal"] = out["shape_equal"] and bool(np.array_equal(res.unique_pix.cpu().numpy(), front["uniq"]))
    n_rng = max(S * P_, 128 * ((out["U"] + 127) // 128))
    states = rng_states(S * P_, rng_seed)
    if n_rng > S * P_:
        states = np.concatenate([states, rng_states(n_rng - S * P_, rng_seed)])
    o_sig = orc.tracks_current_mc(otr, front["neigh"], front["T"], response, states, 0)
    g_sig = res.signals.cpu().numpy()
    out["signals_relerr"] = rel_err(g_sig, o_sig)
    out["signals_sum"] = (float(g_sig.astype(np.float64).sum()), float(o_sig.astype(np.float64).sum()))
    back = oracle_back(orc, front, g_sig, states)
    out["tpm_equal"] = bool(np.array_equal(res.track_pixel_map.cpu().numpy(), back["tpm"]))
    out["pixels_signals_equal"] = bool(np.array_equal(res.pixels_signals.cpu().numpy(), back["ps"]))
    g_digit = res.adc_digit.cpu().numpy()
    out["adc_mismatch"] = int((g_digit != back["digit"]).sum())
    g_adc = res.adc_list.cpu().numpy()
    out["adc_list_equal"] = bool(np.array_equal(g_adc, back["adc"]))
    out["adc_list_relerr"] = float(np.abs(g_adc - back["adc"]).max() / max(np.abs(back["adc"]).max(), 1e-300))
    out["adc_pattern_equal"] = bool(np.array_equal(g_adc != 0, back["adc"] != 0))
    out["ticks_equal"] = bool(np.array_equal(res.adc_ticks_list.cpu().numpy(), back["ticks"]))
    g_cf = res.current_fractions.cpu().numpy()
    out["cf_equal"] = bool(np.array_equal(g_cf, back["cf"]))
    out["cf_maxdiff"] = float(np.abs(g_cf - back["cf"]).max()) if g_cf.size else 0.0
    # order-free sums: absolute agreement at 1e-12 of the largest entry (entries of bipolar neighbours cancel to ~0)
    out["cf_close"] = bool(np.allclose(g_cf, back["cf"], rtol=1e-12, atol=1e-12 * max(1.0, float(np.abs(back["cf"]).max()))))
    out["n_hits_oracle"] = int((back["digit"] > orc.digitize(np.zeros(1))[0]).sum())
    ch.close()
    return out
