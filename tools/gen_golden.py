"""Generate golden input/output vectors from the UNMODIFIED reference (build container only).

The reference kernels' own source is compiled for the host by ``numba.njit`` through
``tools/refharness.py`` (Numba's type inference is shared by its CPU and CUDA targets, so this is the
compiled semantics of the reference: float64 promotion through Python-float globals, float32 where both
operands are float32, int64 ``round``), threads executed one at a time in grid order.  The RNG vectors
come from the installed numba's own ``numba.cuda.random`` host functions.

Outputs: small ``.npz`` fixtures under ``tests/golden/`` (committed).  ``tests/test_oracle_golden.py``
pins the C oracle to them on the CPU, ``tests/test_gpu_golden.py`` pins the CUDA kernels on the GPU.
The reference tree does not exist on the GPU box; nothing under tests/ reads /root/reference.

Run:  python tools/gen_golden.py
"""
import importlib
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import refharness as rh  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
KMODS = ("pixels_from_track", "detsim", "fee", "lightLUT", "light_sim", "quenching", "drifting")

mods = rh.load_reference(simulator=False)
from larndsim_b200 import synth  # noqa: E402


def reload_kernels():
    """What cli/simulate_pixels.py:459-464 does after constants change: re-import the kernel modules so
    the JIT sees the new module globals."""
    for name in KMODS:
        mods[name] = importlib.reload(mods[name])
    rh._cache.clear()


def hk(module, name):
    return rh.host_kernel(mods[module], name)


def consts_dict(consts):
    d, s, l, p = consts.detector, consts.sim, consts.light, consts.physics
    out = {}
    for k in ("TIME_SAMPLING", "TIME_PADDING", "TIME_WINDOW", "RESPONSE_SAMPLING", "RESPONSE_BIN_SIZE", "PIXEL_PITCH",
              "SAMPLED_POINTS", "RESET_NOISE_CHARGE", "UNCORRELATED_NOISE_CHARGE", "DISCRIMINATOR_NOISE",
              "DISCRIMINATION_THRESHOLD", "BUFFER_RISETIME", "GAIN", "V_CM", "V_REF", "V_PEDESTAL", "ADC_COUNTS",
              "V_DRIFT", "ELECTRON_LIFETIME", "LONG_DIFF", "TRAN_DIFF", "E_FIELD", "LAR_DENSITY", "CLOCK_CYCLE",
              "ADC_HOLD_DELAY", "ADC_BUSY_DELAY", "RESET_CYCLES"):
        out["detector." + k] = getattr(d, k)
    for k in ("MAX_TRACKS_PER_PIXEL", "MIN_STEP_SIZE", "MC_SAMPLE_MULTIPLIER", "MAX_ADC_VALUES", "MC_TRUTH_THRESHOLD"):
        out["sim." + k] = getattr(s, k)
    for k in ("ENABLE_LUT_SMEARING", "LIGHT_TICK_SIZE", "SINGLET_FRACTION", "TAU_S", "TAU_T", "SIPM_RESPONSE_MODEL",
              "LIGHT_TRIG_MODE", "N_OP_CHANNEL"):
        if hasattr(l, k):
            out["light." + k] = getattr(l, k)
    if hasattr(l, "LIGHT_WINDOW"):
        out["light.LIGHT_WINDOW"] = np.asarray(l.LIGHT_WINDOW, dtype=np.float64)
    return out


def save(name, consts, **arrays):
    os.makedirs(OUT, exist_ok=True)
    meta = {"c:" + k: np.asarray(v) for k, v in consts_dict(consts).items()}
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **meta, **arrays)
    print("wrote %-28s %7.1f KB" % (name + ".npz", os.path.getsize(path) / 1024), flush=True)


def b1_tracks(dtype=synth.segment_dtype):
    starts = np.array([(-10, -20, -20), (-9, -19.2, -19.5), (5, 5, 10)], dtype=float)
    ends = np.array([(-9, -19.2, -19.5), (-7.7, -19, -19), (5.3, 6.5, 10.2)], dtype=float)
    return synth._fill(np.arange(3), starts, ends, 2.1, 0.0, 0, dtype)


# ------------------------------------------------------------------------------------------
def gen_rng():
    from numba import njit
    from numba.cuda.random import (create_xoroshiro128p_states, xoroshiro128p_uniform_float32,
                                   xoroshiro128p_normal_float32, init_xoroshiro128p_states)
    out = {}
    for seed, n, start in ((1, 5, 0), (12345, 4, 7), (2 ** 40 + 3, 3, 0)):
        st = np.zeros(n, dtype=np.dtype([("s0", np.uint64), ("s1", np.uint64)], align=True))
        init_xoroshiro128p_states.py_func if False else None
        from numba.cuda.random import init_xoroshiro128p_states_cpu
        init_xoroshiro128p_states_cpu(st, seed, start)
        out["states_seed%d_n%d_start%d" % (seed, n, start)] = st.view(np.uint64).reshape(n, 2).copy()

    @njit
    def draw(states, n):
        u = np.zeros(n, dtype=np.float32)
        g = np.zeros(n, dtype=np.float32)
        for i in range(n):
            g[i] = xoroshiro128p_normal_float32(states, 0)
        for i in range(n):
            u[i] = xoroshiro128p_uniform_float32(states, 1)
        return u, g
    st = np.zeros(3, dtype=np.dtype([("s0", np.uint64), ("s1", np.uint64)], align=True))
    from numba.cuda.random import init_xoroshiro128p_states_cpu
    init_xoroshiro128p_states_cpu(st, 1, 0)
    u, g = draw(st, 64)
    out["uniform_state1"] = u
    out["normal_state0"] = g
    out["states_after"] = st.view(np.uint64).reshape(3, 2).copy()
    os.makedirs(OUT, exist_ok=True)
    np.savez_compressed(os.path.join(OUT, "rng.npz"), **out)
    print("wrote rng.npz", flush=True)


def gen_segments(tag, detprop, layout, simprop, i_module=-1):
    consts = rh.load_properties(detprop, layout, simprop, i_module)
    reload_kernels()
    det = consts.detector
    phys = consts.physics
    cases = {
        "b1": b1_tracks() if tag == "module0" else synth.cosmic_segments(3, det, seed=99),
        "cosmic": synth.cosmic_segments(60, det, seed=2),
        "beam": synth.beam_spill_segments(60, det, seed=4),
        "f8": synth.cosmic_segments(30, det, seed=6, dtype=synth.test_dtype_f8),
    }
    cases["cosmic"]["dEdx"][:2] = [0.0, 1e10]         # tests/testQuenching.py edge cases
    arrays = {}
    for cname, tr0 in cases.items():
        arrays[cname + ":in"] = tr0.copy()
        box = tr0.copy()
        hk("quenching", "quench")((len(box),), box, phys.BOX)
        arrays[cname + ":quench_box"] = box
        tr = tr0.copy()
        hk("quenching", "quench")((len(tr),), tr, phys.BIRKS)
        arrays[cname + ":quench_birks"] = tr.copy()
        hk("drifting", "drift")((len(tr),), tr)
        arrays[cname + ":drift"] = tr.copy()
        if "pixel_plane" not in tr.dtype.names:
            continue
        S = len(tr)
        mp = np.zeros(1, dtype=np.int64)
        hk("pixels_from_track", "max_pixels")((S,), tr, mp)
        radius = int(np.ceil(max(tr["tran_diff"]) * 5 / det.PIXEL_PITCH))
        P = (2 * radius + 1) * int(mp[0]) + (1 + 2 * radius) * radius * 2
        act = np.full((S, int(mp[0])), -1, dtype=np.int32)
        nb = np.full((S, P), -1, dtype=np.int32)
        nr = np.full((S, P), -1, dtype=np.int32)
        npl = np.zeros(S)
        hk("pixels_from_track", "get_pixels")((S,), tr, act, nb, nr, npl, radius)
        ts = np.zeros(S)
        tm = np.zeros(1, dtype=np.int64)
        hk("detsim", "time_intervals")((S,), ts, tm, tr)
        # glue exactly as cli/simulate_pixels.py:953-956, 1021-1025 (numpy standing in for cupy)
        uniq = np.unique(nb.reshape(-1))
        uniq = uniq[uniq != -1]
        pim = np.full(nb.shape, -1)
        for i_ in range(S):
            compare = nb[i_, ..., np.newaxis] == uniq
            idx = np.where(compare)
            pim[i_, idx[0]] = idx[1]
        arrays.update({cname + ":max_pixels": mp, cname + ":radius": np.array(radius), cname + ":active": act,
                       cname + ":neigh": nb, cname + ":nrad": nr, cname + ":npl": npl, cname + ":starts": ts,
                       cname + ":tmax": tm, cname + ":uniq": uniq.astype(np.int32), cname + ":pim": pim.astype(np.int64)})
        for K in (50, 2):
            tpm = np.full((len(uniq), K), -1, dtype=np.int64)
            hk("detsim", "get_track_pixel_map2")((len(uniq),), tpm, uniq.astype(np.int32), nb, nr, int(nr.max()) + 1)
            arrays[cname + ":tpm2_K%d" % K] = tpm
            tpm = np.full((len(uniq), K), -1, dtype=np.int64)
            hk("detsim", "get_track_pixel_map")((len(uniq),), tpm, uniq.astype(np.int32), nb)
            arrays[cname + ":tpm1_K%d" % K] = tpm
    arrays["digitize:in"] = np.array([0, 3e3, 7e3, 2e4, 1e5, 3e5, -5e3, 1234.5])
    arrays["digitize:out"] = mods["fee"].digitize(arrays["digitize:in"], det.GAIN * consts.units.mV / consts.units.e)
    save("segments_" + tag, consts, **arrays)


def gen_current_and_fee():
    """SURVEY appendix B.2 setup (short time window so the vectors stay small)."""
    consts = rh.load_properties()
    det, sim = consts.detector, consts.sim
    det.TIME_PADDING = 10
    det.TIME_WINDOW = 8.9
    sim.MIN_STEP_SIZE = 0.05
    sim.MAX_TRACKS_PER_PIXEL = 4
    det.SAMPLED_POINTS = 6
    reload_kernels()
    tr = b1_tracks()[:2].copy()
    hk("quenching", "quench")((2,), tr, consts.physics.BIRKS)
    hk("drifting", "drift")((2,), tr)
    S = 2
    mp = np.zeros(1, dtype=np.int64)
    hk("pixels_from_track", "max_pixels")((S,), tr, mp)
    radius = 1
    P = 3 * int(mp[0]) + 6
    act = np.full((S, int(mp[0])), -1, dtype=np.int32)
    nb = np.full((S, P), -1, dtype=np.int32)
    nr = np.full((S, P), -1, dtype=np.int32)
    npl = np.zeros(S)
    hk("pixels_from_track", "get_pixels")((S,), tr, act, nb, nr, npl, radius)
    ts = np.zeros(S)
    tm = np.zeros(1, dtype=np.int64)
    hk("detsim", "time_intervals")((S,), ts, tm, tr)
    T = int(tm[0])
    k = np.arange(90)
    g = (k / 89.0) ** 4
    g = g / (0.1 * g.sum())
    ii, jj = np.meshgrid(np.arange(45), np.arange(45), indexing="ij")
    lut = (np.exp(-(ii ** 2 + jj ** 2) / 50.0)[:, :, None] * g[None, None, :]).astype(np.float32)
    uniq = np.unique(nb.reshape(-1))
    uniq = uniq[uniq != -1].astype(np.int32)
    U = len(uniq)
    pim = np.full(nb.shape, -1)
    for i_ in range(S):
        idx = np.where(nb[i_, ..., np.newaxis] == uniq)
        pim[i_, idx[0]] = idx[1]
    pim = pim.astype(np.int64)
    tpm = np.full((U, sim.MAX_TRACKS_PER_PIXEL), -1, dtype=np.int64)
    hk("detsim", "get_track_pixel_map2")((U,), tpm, uniq, nb, nr, int(nr.max()) + 1)
    from numba.cuda.random import init_xoroshiro128p_states_cpu
    sdt = np.dtype([("s0", np.uint64), ("s1", np.uint64)], align=True)
    arrays = dict(tracks=tr, neigh=nb, nrad=nr, starts=ts, tmax=tm, lut=lut, uniq=uniq, pim=pim, tpm=tpm)
    Tt = len(det.TIME_TICKS)
    time_ticks = np.linspace(0, 200, Tt + 1)
    for label, sigma0 in (("sigma", False), ("sigma0", True)):
        t2 = tr.copy()
        if sigma0:
            t2["tran_diff"] = 0
            t2["long_diff"] = 0
        st = np.zeros(S * P, dtype=sdt)
        init_xoroshiro128p_states_cpu(st, 1, 0)
        sig = np.zeros((S, P, T), dtype=np.float32)
        hk("detsim", "tracks_current_mc")((S, P, T), sig, nb, t2, lut, st)
        arrays["mc_%s:tracks" % label] = t2
        arrays["mc_%s:signals" % label] = sig
        arrays["mc_%s:states_after" % label] = st.view(np.uint64).reshape(-1, 2).copy()
        ps = np.zeros((U, Tt))
        pts = np.zeros((U, Tt, sim.MAX_TRACKS_PER_PIXEL))
        of = np.zeros(U)
        hk("detsim", "sum_pixel_signals")((S, P, T), ps, sig, ts, pim, tpm, pts, of)
        arrays["sum_%s:ps" % label] = ps
        arrays["sum_%s:pts_nonzero_idx" % label] = np.argwhere(pts != 0).astype(np.int32)
        arrays["sum_%s:pts_nonzero_val" % label] = pts[pts != 0]
        arrays["sum_%s:overflow" % label] = of
        for noise in (False, True):
            if not noise:
                saved = (det.RESET_NOISE_CHARGE, det.UNCORRELATED_NOISE_CHARGE, det.DISCRIMINATOR_NOISE)
                det.RESET_NOISE_CHARGE = det.UNCORRELATED_NOISE_CHARGE = det.DISCRIMINATOR_NOISE = 0
                reload_kernels()
            st2 = np.zeros(U, dtype=sdt)
            init_xoroshiro128p_states_cpu(st2, 2, 0)
            adc = np.zeros((U, sim.MAX_ADC_VALUES))
            ticks = np.zeros((U, sim.MAX_ADC_VALUES))
            cf = np.zeros((U, sim.MAX_ADC_VALUES, sim.MAX_TRACKS_PER_PIXEL))
            thr = np.full(U, det.DISCRIMINATION_THRESHOLD * consts.units.e)
            hk("fee", "get_adc_values")((U,), ps, pts, time_ticks, adc, ticks, 0, st2, cf, thr)
            key = "fee_%s_%s" % (label, "noise" if noise else "quiet")
            arrays[key + ":adc"] = adc
            arrays[key + ":ticks"] = ticks
            arrays[key + ":cf"] = cf
            arrays[key + ":states_after"] = st2.view(np.uint64).reshape(-1, 2).copy()
            arrays[key + ":digit"] = mods["fee"].digitize(adc, det.GAIN * consts.units.mV / consts.units.e)
            if not noise:
                det.RESET_NOISE_CHARGE, det.UNCORRELATED_NOISE_CHARGE, det.DISCRIMINATOR_NOISE = saved
                reload_kernels()
    # deterministic tracks_current on a 6x6 grid, first segment only
    sig = np.zeros((1, P, T), dtype=np.float32)
    hk("detsim", "tracks_current")((1, P, T), sig, nb[:1], tr[:1], lut)
    arrays["tc:signals"] = sig
    arrays["time_ticks"] = time_ticks
    save("current_fee_module0", consts, **arrays)


def gen_light():
    """SURVEY appendix B.3 setup."""
    consts = rh.load_properties()
    light, sim = consts.light, consts.sim
    light.ENABLE_LUT_SMEARING = True
    light.LIGHT_WINDOW = (0.05, 0.15)
    for n_true in (0, 2):
        sim.MAX_MC_TRUTH_IDS = n_true
        sim.MC_TRUTH_THRESHOLD = 0.1
        reload_kernels()
        tr = b1_tracks()
        tr["segment_id"] = [100, 101, 102]
        for k in ("t0", "t0_start", "t0_end"):
            tr[k] = [0.010, 0.012, 0.020]
        hk("quenching", "quench")((3,), tr, consts.physics.BIRKS)
        hk("drifting", "drift")((3,), tr)
        lut = synth.light_lut((14, 26, 8, 48), 16)
        ndet = light.N_OP_CHANNEL
        linc = np.zeros((3, ndet), dtype=[("segment_id", "u4"), ("n_photons_det", "f4"), ("t0_det", "f4")])
        vox = np.zeros((3, 3), dtype=np.int32)
        hk("lightLUT", "calculate_light_incidence")((3,), tr, lut, linc, vox)
        nticks, t_start = mods["light_sim"].get_nticks(linc)
        op_channel = np.arange(8, dtype=np.int32)
        nd = len(op_channel)
        sorted_idx = np.zeros((nd, 3), dtype=np.int64)
        for i, ch in enumerate(op_channel):
            sorted_idx[i] = np.argsort(linc["n_photons_det"][:, ch])[::-1]
        inc = np.zeros((nd, nticks), dtype=np.float32)
        tid = np.full((nd, nticks, n_true), -1, dtype=np.int64)
        tph = np.zeros((nd, nticks, n_true), dtype=np.float64)
        seg_ids = tr["segment_id"].astype(np.int64)
        hk("light_sim", "sum_light_signals")((nd, nticks), tr, vox, seg_ids, linc, op_channel, lut, float(t_start), inc, tid,
                                             tph, sorted_idx, float(lut["time_dist"].shape[-1]))
        sc = np.zeros_like(inc)
        sid = np.full_like(tid, -1)
        sph = np.zeros_like(tph)
        hk("light_sim", "calc_scintillation_effect")((nd, nticks), inc, tid, tph, sc, sid, sph)
        from numba.cuda.random import init_xoroshiro128p_states_cpu
        sdt = np.dtype([("s0", np.uint64), ("s1", np.uint64)], align=True)
        st = np.zeros(nd * nticks, dtype=sdt)
        init_xoroshiro128p_states_cpu(st, 3, 0)
        disc = np.zeros_like(inc)
        hk("light_sim", "calc_stat_fluctuations")((nd, nticks), sc, disc, st)
        resp = np.zeros_like(inc)
        rid = np.full_like(tid, -1)
        rph = np.zeros_like(tph)
        hk("light_sim", "calc_light_detector_response")((nd, nticks), sc, sid, sph, resp, rid, rph)
        save("light_module0_true%d" % n_true, consts, tracks=tr, lut=lut, linc=linc, voxel=vox, nticks=np.array(nticks),
             t_start=np.array(t_start, dtype=np.float64), op_channel=op_channel, sorted_idx=sorted_idx, seg_ids=seg_ids,
             inc=inc, inc_id=tid, inc_ph=tph, scint=sc, scint_id=sid, scint_ph=sph, disc=disc,
             states_after=st.view(np.uint64).reshape(-1, 2).copy(), resp=resp, resp_id=rid, resp_ph=rph,
             light_gain=np.asarray(light.LIGHT_GAIN, dtype=np.float64).reshape(-1),
             impulse=np.asarray(light.IMPULSE_MODEL, dtype=np.float64),
             op_eff=np.asarray(light.OP_CHANNEL_EFFICIENCY, dtype=np.float64),
             op_tpc=np.asarray(light.OP_CHANNEL_TO_TPC, dtype=np.int64),
             impulse_tick=np.array(light.IMPULSE_TICK_SIZE), response_time=np.array(light.LIGHT_RESPONSE_TIME),
             osc_period=np.array(light.LIGHT_OSCILLATION_PERIOD))


if __name__ == "__main__":
    which = sys.argv[1:] or ["rng", "segments", "current", "light"]
    if "rng" in which:
        gen_rng()
    if "segments" in which:
        # one fresh process per configuration: larndsim.consts module globals persist across load_properties calls
        import subprocess
        for tag in ("module0", "2x2", "ndlar"):
            subprocess.check_call([sys.executable, os.path.abspath(__file__), "segments:" + tag])
    if "segments:module0" in which:
        gen_segments("module0", "module0.yaml", "multi_tile_layout-2.3.16.yaml", "singles_sim.yaml")
    if "segments:2x2" in which:
        gen_segments("2x2", "2x2.yaml", "multi_tile_layout-2.5.16.yaml", "singles_sim.yaml", 3)
    if "segments:ndlar" in which:
        gen_segments("ndlar", "ndlar-module.yaml", "multi_tile_layout-3.0.40.yaml", "singles_sim_ndlar.yaml")
    if "current" in which:
        gen_current_and_fee()
    if "light" in which:
        gen_light()
