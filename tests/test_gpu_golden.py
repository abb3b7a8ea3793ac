"""CUDA kernels (through the drop-in modules / C ABI) against the golden vectors produced by the
reference's own kernels (tests/golden/*.npz, tools/gen_golden.py).  Integer outputs bit for bit;
float32 outputs identical except where a libm transcendental is involved (stated per assert)."""
import numpy as np
import pytest

import helpers as h
from larndsim_b200 import _launch as ll
from test_oracle_golden import load, SNAPSHOT, _light_setup

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tag", ["module0", "2x2", "ndlar"])
def test_segment_kernels_golden(cuda, tag):
    from larndsim_b200 import quenching, drifting, pixels_from_track, detsim, fee
    g, mod = load("segments_" + tag, SNAPSHOT[tag])
    for case in ("b1", "cosmic", "beam", "f8"):
        tr = g[case + ":in"].copy()
        box = tr.copy()
        quenching.quench[1, 256](box, mod.physics.BOX)
        assert h.records_equal(box, g[case + ":quench_box"], f64_rtol=h.F64_RTOL)
        quenching.quench[1, 256](tr, mod.physics.BIRKS)
        assert h.records_equal(tr, g[case + ":quench_birks"], f64_rtol=h.F64_RTOL)
        drifting.drift[1, 256](tr)
        assert h.records_equal(tr, g[case + ":drift"], f64_rtol=h.F64_RTOL)
        if case + ":neigh" not in g.files:
            continue
        tr = g[case + ":drift"].copy()
        mp = np.array([0])
        pixels_from_track.max_pixels[1, 128](tr, mp)
        assert mp[0] == g[case + ":max_pixels"][0]
        act = np.full_like(g[case + ":active"], -1)
        nb = np.full_like(g[case + ":neigh"], -1)
        nr = np.full_like(nb, -1)
        npl = np.zeros(len(tr))
        pixels_from_track.get_pixels[1, 128](tr, act, nb, nr, npl, int(g[case + ":radius"]))
        assert np.array_equal(act, g[case + ":active"]) and np.array_equal(nb, g[case + ":neigh"])
        assert np.array_equal(nr, g[case + ":nrad"]) and np.array_equal(npl, g[case + ":npl"])
        ts = np.zeros(len(tr))
        tm = np.zeros(1, dtype=np.int64)
        detsim.time_intervals[1, 128](ts, tm, tr)
        assert np.array_equal(ts, g[case + ":starts"]) and tm[0] == g[case + ":tmax"][0]
        uniq, pim = detsim.unique_pixels(nb)
        assert np.array_equal(uniq.cpu().numpy(), g[case + ":uniq"]) and np.array_equal(pim.cpu().numpy(), g[case + ":pim"])
        for K in (50, 2):
            tpm = np.full((len(uniq), K), -1, dtype=np.int64)
            detsim.get_track_pixel_map2[1, 32](tpm, g[case + ":uniq"], nb, nr, int(nr.max()) + 1)
            assert np.array_equal(tpm, g[case + ":tpm2_K%d" % K])
            tpm = np.full((len(uniq), K), -1, dtype=np.int64)
            detsim.get_track_pixel_map[1, 32](tpm, g[case + ":uniq"], nb)
            assert np.array_equal(tpm, g[case + ":tpm1_K%d" % K])
    assert np.array_equal(fee.digitize(g["digitize:in"]), g["digitize:out"])


@pytest.mark.parametrize("label", ["sigma", "sigma0"])
def test_current_sum_fee_golden(cuda, label):
    from larndsim_b200 import detsim, fee, rng
    g, mod = load("current_fee_module0", "module0")
    tr, nb, lut = g["mc_%s:tracks" % label], g["neigh"], g["lut"]
    S, P_ = nb.shape
    ref = g["mc_%s:signals" % label]
    T = ref.shape[2]
    for mode in ("replay", "cloud") if label == "sigma0" else ("replay",):
        detsim.MC_MODE = mode
        try:
            sig = np.zeros((S, P_, T), dtype=np.float32)
            states = rng.create_xoroshiro128p_states(S * P_, 1)
            detsim.tracks_current_mc[(S, P_, T), (1, 1, 1)](sig, nb, tr, lut, states)
        finally:
            detsim.MC_MODE = "cloud"
        assert np.array_equal(sig != 0, ref != 0)
        if label == "sigma0":
            assert h.rel_err(sig, ref) < 1e-6            # float32 group sums vs the reference's float64 accumulation
        else:
            assert h.rel_err(sig, ref) < 1e-5            # + float32 Box-Muller normals (libdevice vs host libm)
        if mode == "replay":
            assert np.array_equal(states.copy_to_host().view(np.uint64).reshape(-1, 2), g["mc_%s:states_after" % label])
    K = int(mod.sim.MAX_TRACKS_PER_PIXEL)
    U, Tt = len(g["uniq"]), len(mod.detector.TIME_TICKS)
    ps = np.zeros((U, Tt)); pts = np.zeros((U, Tt, K)); of = np.zeros(U)
    detsim.sum_pixel_signals[(S, P_, T), (1, 1, 1)](ps, ref, g["starts"], g["pim"], g["tpm"], pts, of)
    assert np.array_equal(ps, g["sum_%s:ps" % label]) and np.array_equal(of, g["sum_%s:overflow" % label])
    assert np.array_equal(np.argwhere(pts != 0), g["sum_%s:pts_nonzero_idx" % label])
    assert np.array_equal(pts[pts != 0], g["sum_%s:pts_nonzero_val" % label])
    A = int(mod.sim.MAX_ADC_VALUES)
    for noise in ("quiet", "noise"):
        saved = (mod.detector.RESET_NOISE_CHARGE, mod.detector.UNCORRELATED_NOISE_CHARGE, mod.detector.DISCRIMINATOR_NOISE)
        if noise == "quiet":
            mod.detector.RESET_NOISE_CHARGE = mod.detector.UNCORRELATED_NOISE_CHARGE = mod.detector.DISCRIMINATOR_NOISE = 0
        adc = np.zeros((U, A)); ticks = np.zeros((U, A)); cf = np.zeros((U, A, K))
        states = rng.create_xoroshiro128p_states(U, 2)
        thr = np.full(U, mod.detector.DISCRIMINATION_THRESHOLD * mod.units.e)
        fee.get_adc_values[1, 128](ps, pts, g["time_ticks"], adc, ticks, 0, states, cf, thr)
        key = "fee_%s_%s" % (label, noise)
        if noise == "quiet":
            assert np.array_equal(adc, g[key + ":adc"])
        else:
            assert np.array_equal(adc != 0, g[key + ":adc"] != 0) and np.allclose(adc, g[key + ":adc"], rtol=1e-7, atol=0)
        assert np.array_equal(ticks, g[key + ":ticks"]) and np.array_equal(cf, g[key + ":cf"])
        assert np.array_equal(states.copy_to_host().view(np.uint64).reshape(-1, 2), g[key + ":states_after"])
        assert np.array_equal(fee.digitize(adc), g[key + ":digit"])
        mod.detector.RESET_NOISE_CHARGE, mod.detector.UNCORRELATED_NOISE_CHARGE, mod.detector.DISCRIMINATOR_NOISE = saved
    ref_tc = g["tc:signals"]
    sig = np.zeros_like(ref_tc)
    detsim.tracks_current[(1, P_, T), (1, 1, 1)](sig, nb[:1], g["tracks"][:1], lut)
    # against the reference's own output: 1e-5 relative wherever the reference's -erf(a) + erf(b) is well-conditioned (the kernel
    # takes the difference through erfc beyond the segment ends, where the reference's form loses its digits: < 1e-6 of the peak)
    assert np.array_equal(sig != 0, ref_tc != 0) and h.rel_err_peak(sig, ref_tc) < 1e-6


@pytest.mark.parametrize("n_true", [0, 2])
def test_light_chain_golden(cuda, n_true):
    from larndsim_b200 import lightLUT, light_sim, rng
    g, mod = _light_setup(n_true)
    tr, lut = g["tracks"], g["lut"]
    ndet = int(mod.light.N_OP_CHANNEL)
    linc = np.zeros((3, ndet), dtype=g["linc"].dtype)
    vox = np.zeros((3, 3), dtype=np.int32)
    lightLUT.calculate_light_incidence[1, 256](tr, lut, linc, vox)
    assert np.array_equal(vox, g["voxel"])
    assert np.array_equal(linc["n_photons_det"], g["linc"]["n_photons_det"]) and np.array_equal(linc["t0_det"], g["linc"]["t0_det"])
    nticks, nd = int(g["nticks"]), len(g["op_channel"])
    inc = np.zeros((nd, nticks), dtype=np.float32)
    tid = np.full((nd, nticks, n_true), -1, dtype=np.int64)
    tph = np.zeros((nd, nticks, n_true))
    light_sim.sum_light_signals[(nd, 4), (1, 64)](tr, vox, g["seg_ids"], g["linc"], g["op_channel"], lut, float(g["t_start"]), inc,
                                                   tid, tph, g["sorted_idx"], float(lut["time_dist"].shape[-1]))
    assert np.array_equal(inc, g["inc"]) and np.array_equal(tid, g["inc_id"]) and np.array_equal(tph, g["inc_ph"])
    sc = np.zeros_like(inc); sid = np.full_like(tid, -1); sph = np.zeros_like(tph)
    light_sim.calc_scintillation_effect[(nd, 4), (1, 64)](g["inc"], g["inc_id"], g["inc_ph"], sc, sid, sph)
    assert np.array_equal(sc, g["scint"]) and np.array_equal(sid, g["scint_id"]) and np.array_equal(sph, g["scint_ph"])
    disc = np.zeros_like(inc)
    states = rng.create_xoroshiro128p_states(nd * nticks, 3)
    light_sim.calc_stat_fluctuations[(nd, 4), (1, 64)](g["scint"], disc, states)
    # Poisson inverse-CDF uses exp(-mean) in float64 and a float32 uniform: counts are integers, identical unless
    # u falls within 1 ulp of a CDF step
    assert np.array_equal(disc, g["disc"])
    assert np.array_equal(states.copy_to_host().view(np.uint64).reshape(-1, 2), g["states_after"])
    resp = np.zeros_like(inc); rid = np.full_like(tid, -1); rph = np.zeros_like(tph)
    light_sim.calc_light_detector_response[(nd, 4), (1, 64)](g["scint"], g["scint_id"], g["scint_ph"], resp, rid, rph)
    assert np.array_equal(resp, g["resp"]) and np.array_equal(rid, g["resp_id"]) and np.array_equal(rph, g["resp_ph"])
