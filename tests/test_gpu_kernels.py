"""Parity of every CUDA kernel with the CPU oracle, called through the drop-in modules (which go through
the C ABI) with the argument kinds the reference CLI uses: host structured arrays and device arrays."""
import ctypes as C

import numpy as np
import pytest

import helpers as h
from larndsim_b200 import consts as lc, synth
from larndsim_b200 import _launch as ll

pytestmark = pytest.mark.gpu

CONFIGS = ["module0", "2x2", "ndlar"]


def _tracks(n, config, seed=3, kind="cosmic", dtype=None):
    return h.production_tracks(n, config, seed, kind, dtype)


def _f8_tracks(n, config, seed=5):
    return h.production_tracks(n, config, seed, "cosmic", synth.test_dtype_f8)


@pytest.mark.parametrize("config", CONFIGS)
@pytest.mark.parametrize("mode", [1, 2])
def test_quench_drift_bitexact(cuda, config, mode):
    from larndsim_b200 import quenching, drifting
    for tr in (_tracks(3000, config), _f8_tracks(500, config), _tracks(2000, config, kind="beam")):
        tr["dEdx"][:3] = [0.0, 1e10, 2.1]
        ref = tr.copy()
        orc = h.Oracle()
        assert orc.quench(ref, mode) == 0
        orc.drift(ref)
        got = tr.copy()
        quenching.quench[(len(got) + 255) // 256, 256](got, mode)
        drifting.drift[(len(got) + 255) // 256, 256](got)
        assert h.records_equal(got, ref, f64_rtol=h.F64_RTOL)     # f4/u4/i4 fields bit-exact
        assert (ref["pixel_plane"] != 0xBEEF).all() or config != "module0"


def test_quench_invalid_mode_raises(cuda):
    from larndsim_b200 import quenching
    lc.load_snapshot("module0")
    with pytest.raises(ValueError):
        quenching.quench[1, 256](_tracks(10, "module0"), 7)


def test_empty_inputs(cuda):
    from larndsim_b200 import quenching, drifting, pixels_from_track
    lc.load_snapshot("module0")
    tr = _tracks(10, "module0")[:0]
    quenching.quench[1, 256](tr, 2)
    drifting.drift[1, 256](tr)
    m = np.array([0])
    pixels_from_track.max_pixels[1, 128](tr, m)
    assert m[0] == 0


@pytest.mark.parametrize("config", CONFIGS)
@pytest.mark.parametrize("kind", ["cosmic", "beam"])
def test_pixels_and_maps_bitexact(cuda, config, kind):
    from larndsim_b200 import quenching, drifting, pixels_from_track, detsim
    import torch
    tr = _tracks(4000, config, kind=kind)
    orc = h.Oracle()
    front = h.oracle_front(tr.copy(), orc)
    quenching.quench[16, 256](tr, 2)
    drifting.drift[16, 256](tr)
    m = np.array([0])
    pixels_from_track.max_pixels[32, 128](tr, m)
    assert m[0] == front["maxpix"]
    S, P_ = front["neigh"].shape
    act = torch.full((S, front["maxpix"]), -1, dtype=torch.int32, device="cuda")
    nb = torch.full((S, P_), -1, dtype=torch.int32, device="cuda")
    nr = torch.full((S, P_), -1, dtype=torch.int32, device="cuda")
    npl = torch.zeros(S, dtype=torch.float64, device="cuda")
    pixels_from_track.get_pixels[32, 128](tr, act, nb, nr, npl, front["radius"])
    assert np.array_equal(act.cpu().numpy(), front["active"])
    assert np.array_equal(nb.cpu().numpy(), front["neigh"])
    assert np.array_equal(nr.cpu().numpy(), front["nrad"])
    assert np.array_equal(npl.cpu().numpy(), front["npl"])
    uniq, pim = detsim.unique_pixels(nb)
    assert np.array_equal(uniq.cpu().numpy(), front["uniq"])
    assert np.array_equal(pim.cpu().numpy(), orc.pixel_index_map(front["neigh"], front["uniq"]))
    ts = torch.empty(S, dtype=torch.float64, device="cuda")
    tm = torch.zeros(1, dtype=torch.int64, device="cuda")
    detsim.time_intervals[32, 128](ts, tm, tr)
    assert np.array_equal(ts.cpu().numpy(), front["starts"]) and int(tm.item()) == front["T"]
    for K in (50, 3):
        tpm = torch.full((len(uniq), K), -1, dtype=torch.int64, device="cuda")
        detsim.get_track_pixel_map2[(len(uniq) + 31) // 32, 32](tpm, uniq, nb, nr, int(nr.max().item()) + 1)
        assert np.array_equal(tpm.cpu().numpy(), orc.track_pixel_map2(front["uniq"], front["neigh"], front["nrad"],
                                                                       int(front["nrad"].max()) + 1, K))
        tpm1 = np.full((len(uniq), K), -1, dtype=np.int64)       # host array: copy-in / copy-out path
        detsim.get_track_pixel_map[(len(uniq) + 31) // 32, 32](tpm1, front["uniq"], front["neigh"])
        assert np.array_equal(tpm1, orc.track_pixel_map(front["uniq"], front["neigh"], K))


def test_track_pixel_map_unsorted_unique_falls_back(cuda):
    from larndsim_b200 import detsim
    lc.load_snapshot("module0")
    tr = _tracks(300, "module0")
    orc = h.Oracle()
    front = h.oracle_front(tr, orc)
    uniq = front["uniq"][::-1].copy()
    tpm = np.full((len(uniq), 50), -1, dtype=np.int64)
    detsim.get_track_pixel_map2[1, 32](tpm, uniq, front["neigh"], front["nrad"], 3)
    assert np.array_equal(tpm, orc.track_pixel_map2(uniq, front["neigh"], front["nrad"], 3, 50))


def _mc_setup(n, config, seed=11, sigma0=False):
    mod = lc.load_snapshot(config)
    tr = _tracks(n, config, seed=seed)
    orc = h.Oracle()
    front = h.oracle_front(tr, orc)
    if sigma0:
        tr["tran_diff"] = 0
        tr["long_diff"] = 0
    resp = synth.response_lut(mod.detector)
    return mod, tr, orc, front, resp


@pytest.mark.parametrize("config", ["module0", "2x2"])
@pytest.mark.parametrize("mode", ["cloud", "replay"])
def test_tracks_current_mc_vs_oracle(cuda, config, mode):
    """Same RNG states in, same stream discipline: waveforms agree to 1e-5 (float32 outputs; the oracle
    accumulates in float64 like the reference, the kernel in float32 partial sums)."""
    from larndsim_b200 import detsim, rng
    import torch
    n = 24 if mode == "cloud" else 3
    mod, tr, orc, front, resp = _mc_setup(n, config)
    S, P_ = front["neigh"].shape
    T = front["T"] if mode == "cloud" else 96
    if mode == "replay":
        # keep the sequential oracle cheap: a short tick window around the arrival time
        mod.detector.TIME_PADDING = 12.0
        mod.detector.TIME_WINDOW = 11.0
        resp = np.ascontiguousarray(resp[:, :, -140:])
        orc = h.Oracle()
    st0 = h.rng_states(S * P_, 1)
    st_ref = st0.copy()
    ref = orc.tracks_current_mc(tr, front["neigh"], T, resp, st_ref, 1 if mode == "replay" else 0)
    detsim.MC_MODE = mode
    try:
        sig = torch.zeros((S, P_, T), dtype=torch.float32, device="cuda")
        states = rng.create_xoroshiro128p_states(S * P_, 1)
        assert np.array_equal(states.copy_to_host().view(np.uint64).reshape(-1, 2), st0)
        detsim.tracks_current_mc[(S, P_, (T + 63) // 64), (1, 1, 64)](sig, front["neigh"], tr, resp, states)
    finally:
        detsim.MC_MODE = "cloud"
    got = sig.cpu().numpy()
    assert (ref != 0).sum() > 100
    assert h.rel_err(got, ref) < 1e-5
    assert np.array_equal((got != 0), (ref != 0))
    assert np.array_equal(states.copy_to_host().view(np.uint64).reshape(-1, 2), st_ref)   # streams advanced identically


def test_tracks_current_mc_sigma0_is_rng_independent(cuda):
    """tran_diff = long_diff = 0: the normals are multiplied by 0, cloud and replay modes must agree."""
    from larndsim_b200 import detsim, rng
    import torch
    mod, tr, orc, front, resp = _mc_setup(6, "module0", sigma0=True)
    S, P_ = front["neigh"].shape
    T = front["T"]
    out = {}
    for mode in ("cloud", "replay"):
        detsim.MC_MODE = mode
        sig = torch.zeros((S, P_, T), dtype=torch.float32, device="cuda")
        detsim.tracks_current_mc[(S, P_, (T + 63) // 64), (1, 1, 64)](sig, front["neigh"], tr, resp,
                                                                     rng.create_xoroshiro128p_states(S * P_, 5))
        out[mode] = sig.cpu().numpy()
    detsim.MC_MODE = "cloud"
    ref = orc.tracks_current_mc(tr, front["neigh"], T, resp, h.rng_states(S * P_, 9), 1)
    assert h.rel_err(out["cloud"], ref) < 1e-5 and h.rel_err(out["replay"], ref) < 1e-5
    # charge conservation: sum(I dt) of the collecting pixels is of the order of the drifted charge
    q = out["cloud"].astype(np.float64).sum() * mod.detector.TIME_SAMPLING
    assert 0.5 * tr["n_electrons"].sum() < q < 1.5 * tr["n_electrons"].sum()


def test_tracks_current_f64_response(cuda):
    from larndsim_b200 import detsim, rng
    import torch
    mod, tr, orc, front, resp = _mc_setup(8, "module0")
    resp64 = resp.astype(np.float64) * (1 + 1e-9)
    S, P_ = front["neigh"].shape
    st = h.rng_states(S * P_, 2)
    ref = orc.tracks_current_mc(tr, front["neigh"], front["T"], resp64, st, 0)
    sig = torch.zeros((S, P_, front["T"]), dtype=torch.float32, device="cuda")
    detsim.tracks_current_mc[(S, P_, 31), (1, 1, 64)](sig, front["neigh"], tr, resp64, rng.create_xoroshiro128p_states(S * P_, 2))
    assert h.rel_err(sig.cpu().numpy(), ref) < 1e-5


@pytest.mark.parametrize("sampled_points,n_seg,tol", [(8, 3, 2.5e-5), (12, 3, 1e-5), (40, 1, 1e-5)])
def test_tracks_current_deterministic_vs_oracle(cuda, sampled_points, n_seg, tol):
    """tracks_current (detsim.py:351-453): 1e-5 relative on the float32 waveforms, at a reduced grid and at the production
    SAMPLED_POINTS = 40 (40 x 40 x z_steps rho evaluations per pair).  rho's erf difference is evaluated through erfc where the
    reference's -erf(a) + erf(b) cancels (erf_diff, in the kernel and in the oracle alike), so the result no longer depends on
    the last bit of the libm in use."""
    from larndsim_b200 import detsim
    import torch
    mod, tr, orc, front, resp = _mc_setup(n_seg, "module0")
    mod.detector.SAMPLED_POINTS = sampled_points
    orc = h.Oracle()
    S, P_ = front["neigh"].shape
    T = front["T"]
    ref = orc.tracks_current(tr, front["neigh"], T, resp)
    sig = torch.zeros((S, P_, T), dtype=torch.float32, device="cuda")
    detsim.tracks_current[(S, P_, (T + 63) // 64), (1, 1, 64)](sig, front["neigh"], tr, resp)
    got = sig.cpu().numpy()
    assert (ref != 0).sum() > 100
    assert np.array_equal(got != 0, ref != 0)
    # allclose(rtol = tol, atol = tol x waveform peak): north_star's 1e-5 at the production grid (and from 12 points up; measured
    # 8.7e-6 at 12 points, 1.8e-5 on the coarse 6- and 8-point grids where the outermost z slabs carry more of the charge; the bound
    # was 1e-4 before the erf fix)
    e_peak, e_floor = h.rel_err_peak(got, ref), h.rel_err(got, ref)
    q_got, q_ref = got.astype(np.float64).sum(axis=-1), ref.astype(np.float64).sum(axis=-1)
    e_q = float(np.abs(q_got - q_ref).max() / np.abs(q_ref).max())
    print("tracks_current SP=%d: rel_err_peak %.3g, rel_err (1%% floor) %.3g, charge %.3g" % (sampled_points, e_peak, e_floor, e_q))
    assert e_peak < tol
    # with a floor of 1% of the peak instead: ticks where the bipolar waveform passes through zero (measured 1.1e-4, i.e. an
    # absolute 1e-7 of the peak -- the float64 sums over ~1e5 grid points x table values are associated differently)
    assert h.rel_err(got, ref) < 5e-4
    assert e_q < 5 * tol                                                # integrated charge per pixel, relative to the largest


@pytest.mark.parametrize("K", [50, 2])
def test_sum_pixel_signals_and_adc_bitexact(cuda, K):
    """sum_pixel_signals + get_adc_values + digitize from identical inputs, noise ON with seed-matched
    streams: float64 sums in the oracle's order, hit pattern / timestamps / fractions / ADC counts bit for
    bit, RNG streams advanced identically.  The integrated charge carries the float32 Box-Muller normals
    (logf/cosf of libdevice vs glibc, last-bit differences) and is compared to 1e-7."""
    from larndsim_b200 import detsim, fee, rng
    import torch
    mod, tr, orc, front, resp = _mc_setup(40, "module0", seed=21)
    mod.sim.MAX_TRACKS_PER_PIXEL = K
    orc = h.Oracle()
    S, P_ = front["neigh"].shape
    st = h.rng_states(S * P_, 4)
    sig = orc.tracks_current_mc(tr, front["neigh"], front["T"], resp, st, 0)
    st_fee = st.copy()
    back = h.oracle_back(orc, front, sig, st_fee)
    U, Tt, A = len(front["uniq"]), orc.c.n_time_ticks, orc.c.max_adc_values
    assert (back["overflow"] != 0).any() == (K == 2)
    ps = torch.zeros((U, Tt), dtype=torch.float64, device="cuda")
    pts = torch.zeros((U, Tt, K), dtype=torch.float64, device="cuda")
    of = torch.zeros(U, dtype=torch.float64, device="cuda")
    detsim.sum_pixel_signals[(S, P_, 31), (1, 1, 64)](ps, sig, front["starts"], back["pim"], back["tpm"], pts, of)
    assert np.array_equal(ps.cpu().numpy(), back["ps"])
    assert np.array_equal(pts.cpu().numpy(), back["pts"])
    assert np.array_equal(of.cpu().numpy(), back["overflow"])
    time_ticks = np.linspace(0, orc.c.time_interval[1], Tt + 1)
    adc = torch.zeros((U, A), dtype=torch.float64, device="cuda")
    tks = torch.zeros((U, A), dtype=torch.float64, device="cuda")
    cf = torch.zeros((U, A, K), dtype=torch.float64, device="cuda")
    thr = torch.full((U,), orc.c.discrimination_threshold * orc.c.unit_e, dtype=torch.float64, device="cuda")
    states = ll.DeviceRecords(host=st.view(rng.xoroshiro128p_dtype).reshape(-1))
    fee.get_adc_values[(U + 127) // 128, 128](ps, pts, time_ticks, adc, tks, 0, states, cf, thr)
    assert (back["adc"] != 0).sum() > 10
    assert np.array_equal(adc.cpu().numpy() != 0, back["adc"] != 0)
    assert np.allclose(adc.cpu().numpy(), back["adc"], rtol=1e-7, atol=0)
    assert np.array_equal(tks.cpu().numpy(), back["ticks"])
    assert np.array_equal(cf.cpu().numpy(), back["cf"])
    assert np.array_equal(states.copy_to_host().view(np.uint64).reshape(-1, 2), st_fee)
    dg = fee.digitize(adc)
    assert np.array_equal(dg.cpu().numpy(), back["digit"])
    assert np.array_equal(fee.digitize(back["adc"]), back["digit"])            # NumPy in -> NumPy out


def test_adc_noise_off_and_low_threshold(cuda):
    """Noise constants 0 (the RNG still advances) and a low threshold so pixels fire several times,
    reset windows and busy ticks are exercised."""
    from larndsim_b200 import fee, rng
    import torch
    mod, tr, orc, front, resp = _mc_setup(30, "module0", seed=33)
    for k in ("RESET_NOISE_CHARGE", "UNCORRELATED_NOISE_CHARGE", "DISCRIMINATOR_NOISE"):
        setattr(mod.detector, k, 0)
    mod.detector.DISCRIMINATION_THRESHOLD = 1500.0
    orc = h.Oracle()
    S, P_ = front["neigh"].shape
    st = h.rng_states(max(S * P_, 128 * ((len(front["uniq"]) + 127) // 128)), 4)
    sig = orc.tracks_current_mc(tr, front["neigh"], front["T"], resp, st, 0)
    st_fee = st.copy()
    back = h.oracle_back(orc, front, sig, st_fee)
    U, Tt, A, K = len(front["uniq"]), orc.c.n_time_ticks, orc.c.max_adc_values, orc.c.max_tracks_per_pixel
    assert ((back["adc"] != 0).sum(axis=1) > 1).any()
    adc = np.zeros((U, A)); tks = np.zeros((U, A)); cf = np.zeros((U, A, K))
    states = ll.DeviceRecords(host=st.view(rng.xoroshiro128p_dtype).reshape(-1))
    fee.get_adc_values[(U + 127) // 128, 128](back["ps"], back["pts"], np.linspace(0, orc.c.time_interval[1], Tt + 1), adc, tks, 0,
                                              states, cf, np.full(U, 1500.0))
    assert np.array_equal(adc, back["adc"]) and np.array_equal(tks, back["ticks"]) and np.array_equal(cf, back["cf"])
    assert np.array_equal(states.copy_to_host().view(np.uint64).reshape(-1, 2), st_fee)


@pytest.mark.parametrize("config,kind,n", [("module0", "cosmic", 200), ("2x2", "beam", 300), ("ndlar", "beam", 200)])
def test_chain_vs_oracle(cuda, config, kind, n):
    # (noise, dense buffers, fractions in the reference's summation order)
    for noise, dense, exact in ((True, False, True), (False, False, True), (True, True, True), (True, False, False)):
        r = h.chain_vs_oracle(n_segments=n, config=config, seed=17, noise=noise, kind=kind, dense=dense, exact_fractions=exact)
        assert r["tracks_equal"] and r["shape_equal"] and r["unique_equal"] and r["tpm_equal"]
        assert r["signals_relerr"] < 1e-5
        assert r["pixels_signals_equal"] and r["ticks_equal"] and r["adc_pattern_equal"]
        assert r["cf_equal"] if exact else r["cf_close"]        # order-free fraction sums: 1e-12, everything else identical
        assert r["adc_list_equal"] if not noise else r["adc_list_relerr"] < 1e-7
        assert r["adc_mismatch"] == 0 and r["n_hits"] == r["n_hits_oracle"] and r["n_hits"] > 0
        assert r["launches"] > 10


def test_chain_generic_mc_path_vs_oracle(cuda):
    """The generic interior of tracks_current_mc (one table word per add; what float64 tables and non-unit sampling
    ratios always use) on the production shape: same waveforms to 1e-5, everything downstream identical."""
    from larndsim_b200 import _launch as ll
    lib = ll.lib()
    was = lib.lsb_mc_get_grouped()
    lib.lsb_mc_set_grouped(0)
    try:
        assert lib.lsb_mc_get_grouped() == 0
        for config, kind, n in (("module0", "cosmic", 200), ("2x2", "beam", 300)):
            r = h.chain_vs_oracle(n_segments=n, config=config, seed=17, noise=True, kind=kind, exact_fractions=True)
            assert r["tracks_equal"] and r["shape_equal"] and r["unique_equal"] and r["tpm_equal"]
            assert r["signals_relerr"] < 1e-5
            assert r["pixels_signals_equal"] and r["ticks_equal"] and r["adc_pattern_equal"] and r["cf_equal"]
            assert r["adc_mismatch"] == 0 and r["n_hits"] == r["n_hits_oracle"] and r["n_hits"] > 0
    finally:
        lib.lsb_mc_set_grouped(was)


@pytest.mark.parametrize("mode", [0, 1])
def test_chain_interior_variants_vs_oracle(cuda, mode):
    """Both interior strategies of the grouped tracks_current_mc path -- 8-word windows per aligned group (0) and one aligned
    float4 per distinct offset (1; the default only for phase-split tables) -- forced on every configuration:
    same waveforms to 1e-5, everything downstream identical."""
    from larndsim_b200 import _launch as ll
    lib = ll.lib()
    lib.lsb_mc_get_aligned.restype = C.c_int32
    was = lib.lsb_mc_get_aligned()
    lib.lsb_mc_set_aligned(C.c_int32(mode))
    try:
        assert lib.lsb_mc_get_aligned() == mode
        for config, kind, n in (("module0", "cosmic", 200), ("2x2", "beam", 300), ("ndlar", "beam", 200)):
            r = h.chain_vs_oracle(n_segments=n, config=config, seed=23, noise=True, kind=kind, exact_fractions=True)
            assert r["tracks_equal"] and r["shape_equal"] and r["unique_equal"] and r["tpm_equal"]
            assert r["signals_relerr"] < 1e-5
            assert r["pixels_signals_equal"] and r["ticks_equal"] and r["adc_pattern_equal"] and r["cf_equal"]
            assert r["adc_mismatch"] == 0 and r["n_hits"] == r["n_hits_oracle"] and r["n_hits"] > 0
    finally:
        lib.lsb_mc_set_aligned(C.c_int32(was))


def test_light_chain_medium_vs_oracle(cuda):
    """Light chain on 300 cosmic segments, 24 channels, ~1.2k ticks, 9000-tap windows shortened to 400 taps so the
    single-threaded oracle finishes in seconds: bit-exact float32 waveforms (reference add order)."""
    from larndsim_b200 import quenching, drifting, lightLUT, light_sim, rng
    mod = lc.load_snapshot("module0")
    li = mod.light
    li.ENABLE_LUT_SMEARING = True
    li.LIGHT_WINDOW = (0.1, 0.5)
    li.LIGHT_TICK_SIZE = 0.001
    tr = _tracks(300, "module0", seed=8)
    tr["t0"] = np.random.default_rng(3).uniform(0, 0.4, len(tr))
    tr["t0_start"] = tr["t0"]; tr["t0_end"] = tr["t0"]
    orc = h.Oracle()
    orc.quench(tr, 2); orc.drift(tr)
    lut = synth.light_lut((14, 26, 8, 48), 16)
    ol = h.OracleLight()
    ndet = int(li.N_OP_CHANNEL)
    eff = np.asarray(li.OP_CHANNEL_EFFICIENCY, dtype=np.float64)
    tpc = np.asarray(li.OP_CHANNEL_TO_TPC, dtype=np.int64)
    linc_ref, vox_ref = ol.light_incidence(tr, lut, ndet, eff, tpc)
    linc = np.zeros_like(linc_ref); vox = np.zeros_like(vox_ref)
    lightLUT.calculate_light_incidence[2, 256](tr, lut, linc, vox)
    assert np.array_equal(vox, vox_ref) and np.array_equal(linc["n_photons_det"], linc_ref["n_photons_det"])
    assert np.array_equal(linc["t0_det"], linc_ref["t0_det"])
    op_channel = np.arange(0, 96, 4, dtype=np.int32)
    nd, nticks = len(op_channel), 1200
    sorted_idx = np.stack([np.argsort(linc["n_photons_det"][:, ch], kind="stable")[::-1] for ch in op_channel]).astype(np.int64)
    seg_ids = np.arange(len(tr), dtype=np.int64)
    empty_i = np.zeros((nd, nticks, 0), dtype=np.int64); empty_f = np.zeros((nd, nticks, 0))
    ref_inc, _, _ = ol.sum_light_signals(tr, vox, seg_ids, linc, op_channel, lut, -0.1, nticks, 0, sorted_idx, 16)
    inc = np.zeros((nd, nticks), dtype=np.float32)
    light_sim.sum_light_signals[(nd, 19), (1, 64)](tr, vox, seg_ids, linc, op_channel, lut, -0.1, inc, empty_i, empty_f, sorted_idx, 16.0)
    assert (ref_inc != 0).sum() > 500 and np.array_equal(inc, ref_inc)
    ref_sc, _, _ = ol.scintillation(ref_inc, empty_i, empty_f)
    sc = np.zeros_like(inc)
    light_sim.calc_scintillation_effect[(nd, 19), (1, 64)](inc, empty_i, empty_f, sc, empty_i.copy(), empty_f.copy())
    assert np.array_equal(sc, ref_sc)
    st = h.rng_states(nd * nticks, 5)
    ref_disc = ol.stat_fluctuations(ref_sc, st)
    disc = np.zeros_like(inc)
    states = rng.create_xoroshiro128p_states(nd * nticks, 5)
    light_sim.calc_stat_fluctuations[(nd, 19), (1, 64)](sc, disc, states)
    # Poisson counts are integers: identical unless the float32 uniform lands within an ulp of a CDF step,
    # or (mean >= 30) the Gaussian branch truncates a float32 normal that differs in the last bit
    assert (disc != ref_disc).mean() < 1e-4
    assert np.array_equal(states.copy_to_host().view(np.uint64).reshape(-1, 2), st)
    gain = np.asarray(li.LIGHT_GAIN, dtype=np.float64).reshape(-1)
    ref_resp, _, _ = ol.detector_response(ref_disc, empty_i, empty_f, gain, np.asarray(li.IMPULSE_MODEL, dtype=np.float64))
    resp = np.zeros_like(inc)
    light_sim.calc_light_detector_response[(nd, 19), (1, 64)](ref_disc, empty_i, empty_f, resp, empty_i.copy(), empty_f.copy())
    assert np.array_equal(resp, ref_resp)


def test_light_chain_2x2_module_shape_full_taps(cuda):
    """Light chain at the 2x2 shape (BASELINE configs[3]): one module's 96 optical channels x 16 000 ticks of 1 ns, beam-spill
    segments, LUT smearing on, the FULL 16 000-tap scintillation / response windows -- float32 waveforms bit for bit
    (reference add order), through the binned sum_light_signals and the FIRs with trimmed zero taps."""
    from larndsim_b200 import lightLUT, light_sim, rng
    mod = lc.load_snapshot("2x2")
    li = mod.light
    li.ENABLE_LUT_SMEARING = True
    assert tuple(li.LIGHT_WINDOW) == (0, 16) and li.LIGHT_TICK_SIZE == 0.001
    tr = synth.beam_spill_segments(1500, mod.detector, seed=8, n_events=1)
    orc = h.Oracle()
    orc.quench(tr, 2); orc.drift(tr)
    lut = synth.light_lut((14, 26, 8, 48), 40)
    ol = h.OracleLight()
    ndet = int(li.N_OP_CHANNEL)
    eff = np.asarray(li.OP_CHANNEL_EFFICIENCY, dtype=np.float64)
    tpc = np.asarray(li.OP_CHANNEL_TO_TPC, dtype=np.int64)
    linc_ref, vox_ref = ol.light_incidence(tr, lut, ndet, eff, tpc)
    linc = np.zeros_like(linc_ref); vox = np.zeros_like(vox_ref)
    lightLUT.calculate_light_incidence[6, 256](tr, lut, linc, vox)
    assert np.array_equal(vox, vox_ref) and np.array_equal(linc["n_photons_det"], linc_ref["n_photons_det"])
    nticks, t_start = light_sim.get_nticks(linc)
    assert nticks == 16000
    op_channel = np.asarray(li.TPC_TO_OP_CHANNEL)[:2].ravel().astype(np.int32)         # the module's 96 channels (mod2mod mode)
    nd = len(op_channel)
    sorted_idx = np.stack([np.argsort(linc["n_photons_det"][:, ch], kind="stable")[::-1] for ch in op_channel]).astype(np.int64)
    seg_ids = np.arange(len(tr), dtype=np.int64)
    empty_i = np.zeros((nd, nticks, 0), dtype=np.int64); empty_f = np.zeros((nd, nticks, 0))
    ref_inc, _, _ = ol.sum_light_signals(tr, vox, seg_ids, linc, op_channel, lut, t_start, nticks, 0, sorted_idx, 40)
    inc = np.zeros((nd, nticks), dtype=np.float32)
    BPG, TPB = (nd, (nticks + 63) // 64), (1, 64)
    light_sim.sum_light_signals[BPG, TPB](tr, vox, seg_ids, linc, op_channel, lut, t_start, inc, empty_i, empty_f, sorted_idx, 40.0)
    assert (ref_inc != 0).sum() > 1000 and np.array_equal(inc, ref_inc)
    ref_sc, _, _ = ol.scintillation(ref_inc, empty_i, empty_f)
    sc = np.zeros_like(inc)
    light_sim.calc_scintillation_effect[BPG, TPB](inc, empty_i, empty_f, sc, empty_i.copy(), empty_f.copy())
    assert np.array_equal(sc, ref_sc)
    st = h.rng_states(nd * nticks, 5)
    ref_disc = ol.stat_fluctuations(ref_sc, st)
    disc = np.zeros_like(inc)
    states = rng.create_xoroshiro128p_states(nd * nticks, 5)
    light_sim.calc_stat_fluctuations[BPG, TPB](sc, disc, states)
    assert (disc != ref_disc).mean() < 1e-4
    assert np.array_equal(states.copy_to_host().view(np.uint64).reshape(-1, 2), st)
    gain = np.asarray(li.LIGHT_GAIN, dtype=np.float64).reshape(-1)
    ref_resp, _, _ = ol.detector_response(ref_disc, empty_i, empty_f, gain, np.asarray(li.IMPULSE_MODEL, dtype=np.float64))
    resp = np.zeros_like(inc)
    light_sim.calc_light_detector_response[BPG, TPB](ref_disc, empty_i, empty_f, resp, empty_i.copy(), empty_f.copy())
    assert (ref_resp != 0).sum() > 1000 and np.array_equal(resp, ref_resp)
