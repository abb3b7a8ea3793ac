"""Pins the CPU oracle (oracle/larnd_oracle.c) to golden vectors produced by the reference's own kernels
(tools/gen_golden.py: reference source compiled for the host by numba, RNG vectors from numba.cuda.random).
CPU only.  Integer / float32 outputs must be identical; float64 values that pass through exp/log are
identical here too because the generator and the oracle both use the host libm."""
import os

import numpy as np
import pytest

import helpers as h
from larndsim_b200 import consts as lc, _abi

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
SNAPSHOT = {"module0": "module0", "2x2": "2x2_mod2mod_variation_mod3", "ndlar": "ndlar"}


def load(name, snapshot):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    mod = lc.load_snapshot(snapshot)
    for k in g.files:
        if k.startswith("c:"):
            ns, attr = k[2:].split(".")
            v = g[k]
            v = tuple(v.tolist()) if v.ndim else v.item()
            setattr(getattr(mod, ns), attr, v)
    return g, mod


def test_rng_matches_numba():
    g = np.load(os.path.join(GOLD, "rng.npz"))
    for k in g.files:
        if k.startswith("states_seed"):
            seed, n, start = (int(x[len(p):]) for x, p in zip(k.split("_")[1:], ("seed", "n", "start")))
            assert np.array_equal(h.rng_states(n, seed, start), g[k]), k
    st = h.rng_states(3, 1)
    # SURVEY appendix B.1
    assert st[0, 0] == 0x910a2dec89025cc1 and st[1, 1] == 0x6c96c932f21d71ee and st[2, 0] == 0x0b3d0cd6ddfa7a4c
    nrm = np.zeros(64, dtype=np.float32)
    uni = np.zeros(64, dtype=np.float32)
    lib = h.oracle_lib()
    lib.orc_rng_draw(h.P(st[0]), 64, 1, h.P(nrm))
    lib.orc_rng_draw(h.P(st[1]), 64, 0, h.P(uni))
    assert np.array_equal(uni, g["uniform_state1"])
    assert np.allclose(nrm, g["normal_state0"], rtol=3e-7, atol=1e-7)       # float32 logf/cosf: numba vs glibc
    assert np.array_equal(st, g["states_after"])
    assert np.allclose(g["normal_state0"][:4], [-1.44706387, -1.66034780, 1.27061935, -1.11583740], rtol=1e-6)
    assert np.allclose(g["uniform_state1"][:3], [0.87634873, 0.21028069, 0.41908127], rtol=1e-7)


@pytest.mark.parametrize("tag", ["module0", "2x2", "ndlar"])
def test_segment_kernels(tag):
    g, mod = load("segments_" + tag, SNAPSHOT[tag])
    orc = h.Oracle()
    c = orc.c
    for case in ("b1", "cosmic", "beam", "f8"):
        tr = g[case + ":in"].copy()
        box = tr.copy()
        assert orc.quench(box, c.mode_box) == 0
        assert h.records_equal(box, g[case + ":quench_box"]), (tag, case, "box")
        assert orc.quench(tr, c.mode_birks) == 0
        assert h.records_equal(tr, g[case + ":quench_birks"]), (tag, case, "birks")
        orc.drift(tr)
        assert h.records_equal(tr, g[case + ":drift"]), (tag, case, "drift")
        if case + ":neigh" not in g.files:
            continue
        assert orc.max_pixels(tr) == int(g[case + ":max_pixels"][0])
        radius = int(g[case + ":radius"])
        nb_ref = g[case + ":neigh"]
        act, nb, nr, npl = orc.get_pixels(tr, g[case + ":active"].shape[1], nb_ref.shape[1], radius)
        assert np.array_equal(act, g[case + ":active"]) and np.array_equal(nb, nb_ref)
        assert np.array_equal(nr, g[case + ":nrad"]) and np.array_equal(npl, g[case + ":npl"])
        ts, tm = orc.time_intervals(tr)
        assert np.array_equal(ts, g[case + ":starts"]) and tm == int(g[case + ":tmax"][0])
        uniq = orc.unique_pixels(nb)
        assert np.array_equal(uniq, g[case + ":uniq"])
        assert np.array_equal(orc.pixel_index_map(nb, uniq), g[case + ":pim"])
        for K in (50, 2):
            assert np.array_equal(orc.track_pixel_map2(uniq, nb, nr, int(nr.max()) + 1, K), g[case + ":tpm2_K%d" % K])
            assert np.array_equal(orc.track_pixel_map(uniq, nb, K), g[case + ":tpm1_K%d" % K])
    assert np.array_equal(orc.digitize(g["digitize:in"]), g["digitize:out"])


def test_appendix_b1_values():
    """SURVEY.md appendix B.1 (pixel ids, distance classes, track starts)."""
    g, mod = load("segments_module0", "module0")
    assert g["b1:quench_birks"]["n_electrons"].tolist() == [85254, 87261, 95670]
    d = g["b1:drift"]
    assert d["pixel_plane"].tolist() == [0, 0, 1] and d["n_electrons"].tolist() == [83088, 84941, 91096]
    assert int(g["b1:max_pixels"][0]) == 5
    assert g["b1:active"].tolist() == [[20207, 20208, 20348, 20349, -1], [20349, 20350, 20351, 20491, 20492],
                                       [67281, 67421, 67561, 67701, -1]]
    assert g["b1:nrad"][0].tolist() == [2, 1, 2, 1, 0, 1, 2, 1, 2, 2, 1, 2, 2, 1, 2, 2, 1, 2, -1, -1, -1]
    assert g["b1:npl"].tolist() == [18, 21, 18] and len(g["b1:uniq"]) == 48
    assert np.allclose(g["b1:starts"], [-124.7, -121.5, -63.3]) and int(g["b1:tmax"][0]) == 1942
    assert g["digitize:out"][:6].tolist() == [74, 77, 81, 94, 175, 255]


def _states(n, seed):
    return h.rng_states(n, seed)


@pytest.mark.parametrize("label", ["sigma", "sigma0"])
def test_current_sum_fee(label):
    g, mod = load("current_fee_module0", "module0")
    orc = h.Oracle()
    tr = g["mc_%s:tracks" % label]
    nb, nr, lut = g["neigh"], g["nrad"], g["lut"]
    S, P_ = nb.shape
    ts, T = orc.time_intervals(g["tracks"])
    assert np.array_equal(ts, g["starts"]) and T == int(g["tmax"][0])
    st = _states(S * P_, 1)
    sig = orc.tracks_current_mc(tr, nb, T, lut, st, 1)                     # replay = the reference's thread order
    ref = g["mc_%s:signals" % label]
    assert (ref != 0).sum() > 1000
    assert np.array_equal(sig != 0, ref != 0)
    assert h.rel_err(sig, ref) < (1e-6 if label == "sigma" else 1e-7)      # float32 normals: numba vs glibc logf/cosf
    assert np.array_equal(st, g["mc_%s:states_after" % label])
    if label == "sigma0":
        assert np.array_equal(sig, ref)
        # SURVEY appendix B.2 -- those figures came from the CUDA *simulator* (NumPy float32 promotion), the
        # fixtures here from the compiled semantics, hence 1e-6 rather than equality
        assert abs(float(ref.astype(np.float64).sum()) / 2994549.6086 - 1) < 1e-6 and int((ref != 0).sum()) == 4641
        assert abs(float(ref[1, 11, 111]) / 11805.332 - 1) < 1e-6
        cloud = orc.tracks_current_mc(tr, nb, T, lut, _states(S * P_, 77), 0)
        assert np.array_equal(cloud, ref)                                   # sigma=0: RNG discipline is irrelevant
    uniq = orc.unique_pixels(nb)
    assert np.array_equal(uniq, g["uniq"])
    pim = orc.pixel_index_map(nb, uniq)
    K = orc.c.max_tracks_per_pixel
    tpm = orc.track_pixel_map2(uniq, nb, nr, int(nr.max()) + 1, K)
    assert np.array_equal(pim, g["pim"]) and np.array_equal(tpm, g["tpm"])
    Tt = orc.c.n_time_ticks
    ps, pts, of = orc.sum_pixel_signals(ref, ts, pim, tpm, Tt)
    assert np.array_equal(ps, g["sum_%s:ps" % label]) and np.array_equal(of, g["sum_%s:overflow" % label])
    idx = g["sum_%s:pts_nonzero_idx" % label]
    assert np.array_equal(np.argwhere(pts != 0), idx)
    assert np.array_equal(pts[pts != 0], g["sum_%s:pts_nonzero_val" % label])
    for noise in ("quiet", "noise"):
        mod2 = lc.provider()
        saved = (mod2.detector.RESET_NOISE_CHARGE, mod2.detector.UNCORRELATED_NOISE_CHARGE, mod2.detector.DISCRIMINATOR_NOISE)
        if noise == "quiet":
            mod2.detector.RESET_NOISE_CHARGE = mod2.detector.UNCORRELATED_NOISE_CHARGE = mod2.detector.DISCRIMINATOR_NOISE = 0
        o2 = h.Oracle()
        st2 = _states(len(uniq), 2)
        thr = np.full(len(uniq), o2.c.discrimination_threshold * o2.c.unit_e)
        adc, ticks, cf = o2.get_adc_values(ps, pts, g["time_ticks"], o2.c.max_adc_values, 0.0, st2, thr)
        key = "fee_%s_%s" % (label, noise)
        assert (g[key + ":adc"] != 0).sum() >= 10
        if noise == "quiet":
            assert np.array_equal(adc, g[key + ":adc"])
        else:
            assert np.array_equal(adc != 0, g[key + ":adc"] != 0) and np.allclose(adc, g[key + ":adc"], rtol=1e-7, atol=0)
        assert np.array_equal(ticks, g[key + ":ticks"])
        assert np.array_equal(cf, g[key + ":cf"])
        assert np.array_equal(st2, g[key + ":states_after"])
        assert np.array_equal(o2.digitize(adc), g[key + ":digit"])
        mod2.detector.RESET_NOISE_CHARGE, mod2.detector.UNCORRELATED_NOISE_CHARGE, mod2.detector.DISCRIMINATOR_NOISE = saved
    if label == "sigma0":
        # appendix B.2 hits: pixel 20349 fires twice
        i = int(np.nonzero(uniq == 20349)[0][0])
        a = g["fee_sigma0_quiet:adc"][i]
        assert np.allclose(a[:2], [24924.462, 11225.002], rtol=1e-6)
        assert g["fee_sigma0_quiet:digit"][i][:2].tolist() == [99, 85]


def test_tracks_current_deterministic():
    g, mod = load("current_fee_module0", "module0")
    orc = h.Oracle()
    ref = g["tc:signals"]
    T = ref.shape[2]
    got = orc.tracks_current(g["tracks"][:1], g["neigh"][:1], T, g["lut"])
    assert (ref != 0).sum() > 100
    assert np.array_equal(got != 0, ref != 0)
    assert h.rel_err_peak(got, ref) < 1e-6


def _light_setup(n_true):
    g, mod = load("light_module0_true%d" % n_true, "module0")
    li = mod.light
    li.IMPULSE_MODEL = g["impulse"]
    li.IMPULSE_TICK_SIZE = float(g["impulse_tick"])
    li.LIGHT_RESPONSE_TIME = float(g["response_time"])
    li.LIGHT_OSCILLATION_PERIOD = float(g["osc_period"])
    li.LIGHT_GAIN = g["light_gain"]
    li.OP_CHANNEL_EFFICIENCY = g["op_eff"]
    li.OP_CHANNEL_TO_TPC = g["op_tpc"]
    mod.sim.MAX_MC_TRUTH_IDS = n_true
    return g, mod


@pytest.mark.parametrize("n_true", [0, 2])
def test_light_chain(n_true):
    g, mod = _light_setup(n_true)
    ol = h.OracleLight()
    tr, lut = g["tracks"], g["lut"]
    linc, vox = ol.light_incidence(tr, lut, int(mod.light.N_OP_CHANNEL), g["op_eff"], g["op_tpc"])
    assert np.array_equal(vox, g["voxel"]) and vox.tolist() == [[4, 12, 2], [5, 12, 2], [5, 7, 5]]      # appendix B.3
    assert np.array_equal(linc["n_photons_det"], g["linc"]["n_photons_det"])
    assert np.array_equal(linc["t0_det"], g["linc"]["t0_det"])
    nticks = int(g["nticks"])
    inc, tid, tph = ol.sum_light_signals(tr, vox, g["seg_ids"], g["linc"], g["op_channel"], lut, float(g["t_start"]), nticks,
                                         n_true, g["sorted_idx"], lut["time_dist"].shape[-1])
    assert (g["inc"] != 0).sum() == 144 and nticks == 212                                                # appendix B.3
    assert np.array_equal(inc, g["inc"]) and np.array_equal(tid, g["inc_id"]) and np.array_equal(tph, g["inc_ph"])
    sc, sid, sph = ol.scintillation(g["inc"], g["inc_id"], g["inc_ph"])
    assert np.array_equal(sc, g["scint"]) and np.array_equal(sid, g["scint_id"]) and np.array_equal(sph, g["scint_ph"])
    st = h.rng_states(inc.size, 3)
    disc = ol.stat_fluctuations(g["scint"], st)
    assert np.array_equal(disc, g["disc"]) and np.array_equal(st, g["states_after"])
    resp, rid, rph = ol.detector_response(g["scint"], g["scint_id"], g["scint_ph"], g["light_gain"], g["impulse"])
    assert np.array_equal(resp, g["resp"]) and np.array_equal(rid, g["resp_id"]) and np.array_equal(rph, g["resp_ph"])


def test_abi_struct_sizes_match_header():
    """The ctypes mirror must agree with the C structs the oracle was compiled against."""
    lib = h.oracle_lib()
    assert lib.orc_abi_version() == 2
    c = lc.snapshot(lc.load_snapshot("module0"))
    assert c.n_tpc == 2 and c.n_pixels[0] == 140 and abs(c.tpc_borders[5] - (-0.15875)) < 1e-9
