"""Light trigger + digitisation (reference light_sim.py:380-619 get_triggers / sim_triggers /
digitize_signal).  CPU: the restatement (oracle/light_trigger_oracle.py) reproduces what the reference's own functions
returned on the committed cases.  GPU: the CUDA stage (C ABI / light_sim drop-ins) equals the restatement and the
reference fixtures bit for bit (zero noise spectrum)."""
import numpy as np
import pytest

import light_trigger_util as ltu
import light_trigger_oracle as lo


def _case(name):
    z = ltu.load(name)
    C = ltu.consts_from_npz(z)
    sig, op, tid, tph = ltu.case_inputs(name, C["OP_CHANNEL_PER_TRIG"], C["N_OP_CHANNEL"])
    assert np.allclose([sig.astype(np.float64).sum(), tph.sum()], z["in_checksum"], rtol=0, atol=0), "inputs differ from the generator's"
    return z, C, sig, op, tid, tph


@pytest.mark.parametrize("name", ltu.CASES)
def test_oracle_reproduces_reference_triggers_and_waveforms(name):
    z, C, sig, op, tid, tph = _case(name)
    thr = ltu.thresholds(C, op)
    for isub in (0, 1):
        trig, chans, kinds = lo.get_triggers(sig, thr, op, isub, C)
        assert np.array_equal(trig, z["trig_idx_%d" % isub]) and np.array_equal(kinds, z["trig_type_%d" % isub])
        assert chans.shape == z["trig_chan_%d" % isub].shape and np.array_equal(chans, z["trig_chan_%d" % isub])
    trig, chans, _ = lo.get_triggers(sig, thr, op, 0, C)
    assert len(trig) >= 1
    d, d_id, d_ph = lo.sim_triggers(sig, op, tid, tph, trig, chans, int(z["digit_samples"]), C)
    assert d.dtype == np.float64 and np.array_equal(d, z["out_digit"])
    assert np.array_equal(d_id, z["out_digit_id"]) and np.array_equal(d_ph, z["out_digit_photons"])
    assert (d != 0).sum() > 1000


def test_oracle_trigger_search_keeps_the_reference_bookkeeping():
    """module0: pulses at 300, 1500 (inside the dead time), 4200, 4300, 8100 -> the reference reports 300 and 4200 and
    then skips past the end because the remaining waveform is cut at an absolute index."""
    z, C, sig, op, tid, tph = _case("module0")
    trig, _, _ = lo.get_triggers(sig, ltu.thresholds(C, op), op, 0, C)
    assert trig.tolist() == [300, 4200]


# ---------------------------------------------------------------------------------------- GPU
def _provider(name):
    from larndsim_b200 import consts as lc
    return lc.load_snapshot("module0" if name.startswith("module0") else "2x2")


@pytest.mark.gpu
@pytest.mark.parametrize("name", ltu.CASES)
def test_gpu_triggers_and_waveforms_identical(cuda, name):
    import torch
    from larndsim_b200 import light_sim, _launch as ll
    z, C, sig, op, tid, tph = _case(name)
    p = _provider(name)
    assert int(p.light.LIGHT_TRIG_MODE) == C["LIGHT_TRIG_MODE"] and np.array_equal(np.asarray(p.light.LIGHT_TRIG_THRESHOLD), C["LIGHT_TRIG_THRESHOLD"])
    thr = ltu.thresholds(C, op)
    launches0 = ll.lib().lsb_launch_count()
    for isub in (0, 1):
        trig, chans, kinds = light_sim.get_triggers(sig if isub else torch.from_numpy(sig).cuda(), thr, op, isub)
        assert np.array_equal(trig, z["trig_idx_%d" % isub]) and np.array_equal(kinds, z["trig_type_%d" % isub])
        assert np.array_equal(chans, z["trig_chan_%d" % isub])
    trig, chans, _ = light_sim.get_triggers(sig, thr, op, 0)
    ns = int(z["digit_samples"])
    d, d_id, d_ph = light_sim.sim_triggers((1, 1, 1), (1, 1, 64), sig, op, tid, tph, trig, chans, ns, np.zeros((C["N_OP_CHANNEL"], 33)))
    assert ll.lib().lsb_launch_count() > launches0
    d, d_id, d_ph = d.cpu().numpy(), d_id.cpu().numpy(), d_ph.cpu().numpy()
    o, o_id, o_ph = lo.sim_triggers(sig, op, tid, tph, trig, chans, ns, C)
    assert np.array_equal(d, o) and np.array_equal(d_id, o_id) and np.array_equal(d_ph, o_ph)
    assert np.array_equal(d, z["out_digit"]) and np.array_equal(d_id, z["out_digit_id"]) and np.array_equal(d_ph, z["out_digit_photons"])


@pytest.mark.gpu
def test_gpu_raw_digitize_kernel_on_padded_arrays(cuda):
    """digitize_signal[...] itself (no padding, no rounding): fed with the padded arrays sim_triggers would build, and
    rounded afterwards, it gives the reference's waveforms."""
    from math import ceil
    from larndsim_b200 import light_sim
    z, C, sig, op, tid, tph = _case("2x2")
    _provider("2x2")
    trig, chans = z["trig_idx_0"], z["trig_chan_0"]
    ns = int(z["digit_samples"])
    pre = int(ceil(C["LIGHT_TRIG_WINDOW"][0] / C["LIGHT_TICK_SIZE"]))
    front = pre - int(trig.min())
    assert front > 0
    psig = np.concatenate([np.zeros((sig.shape[0], front)), sig], axis=-1)
    ptid = np.concatenate([np.full((sig.shape[0], front, tid.shape[2]), -1, dtype=tid.dtype), tid], axis=1)
    ptph = np.concatenate([np.zeros((sig.shape[0], front, tph.shape[2])), tph], axis=1)
    d = np.zeros((len(trig), chans.shape[1], ns)); d_id = np.full(d.shape + (tid.shape[2],), -1, dtype=np.int64); d_ph = np.zeros(d.shape + (tid.shape[2],))
    light_sim.digitize_signal[(1, 1, 16), (1, 1, 64)](psig, op, trig + front, chans, ptid, ptph, d, d_id, d_ph)
    q = 2 ** (16 - C["LIGHT_NBIT"])
    assert np.array_equal(np.round(d / q) * q, z["out_digit"]) and np.array_equal(d_id, z["out_digit_id"]) and np.array_equal(d_ph, z["out_digit_photons"])


@pytest.mark.gpu
def test_gpu_trigger_edge_cases_and_noise(cuda):
    import torch
    from larndsim_b200 import light_sim
    z, C, sig, op, tid, tph = _case("module0")
    p = _provider("module0")
    thr = ltu.thresholds(C, op)
    # nothing above threshold: no triggers, empty outputs of the reference's shapes
    quiet = np.zeros_like(sig)
    trig, chans, kinds = light_sim.get_triggers(quiet, thr, op, 0)
    assert trig.shape == (0,) and chans.shape == (0, len(op)) and kinds.shape == (0,)
    d, d_id, d_ph = light_sim.sim_triggers((1, 1, 1), (1, 1, 64), sig, op, tid, tph, trig, chans, 256, np.zeros((96, 33)))
    assert tuple(d.shape) == (0, len(op), 256) and tuple(d_id.shape) == (0, len(op), 256, 2)
    # tick count divisible by the sample factor (a whole block of padding) and a pulse at the very end
    s2 = sig[:, :8990].copy()
    o_trig, o_ch, _ = lo.get_triggers(s2, thr, op, 0, C)
    g_trig, g_ch, _ = light_sim.get_triggers(s2, thr, op, 0)
    assert np.array_equal(o_trig, g_trig) and np.array_equal(o_ch, g_ch)
    # noise: flat spectrum -> waveforms differ from the noiseless ones by multiples of the ADC quantum, zero mean, not constant
    trig, chans, _ = light_sim.get_triggers(sig, thr, op, 0)
    spec = np.full((96, 33), 40.0)
    a = light_sim.sim_triggers((1, 1, 1), (1, 1, 64), sig, op, tid, tph, trig, chans, 256, spec)[0].cpu().numpy()
    b = light_sim.sim_triggers((1, 1, 1), (1, 1, 64), sig, op, tid, tph, trig, chans, 256, spec)[0].cpu().numpy()
    clean = z["out_digit"]
    q = 2 ** (16 - C["LIGHT_NBIT"])
    assert np.array_equal(np.round(a / q) * q, a) and (a != clean).mean() > 0.2 and (a != b).mean() > 0.2
    assert abs((a - clean).mean()) < 0.2 * (a - clean).std()
    n = light_sim.gen_light_detector_noise((4, 1001), spec[:4]).cpu().numpy()
    assert n.shape == (4, 1001) and np.array_equal(np.round(n / q) * q, n) and n.std() > 0


@pytest.mark.gpu
def test_gpu_threshold_triggers_across_modules(cuda):
    """Threshold mode on a 4-module geometry (2x2 tables, LIGHT_TRIG_MODE forced to 0): every module searches its own
    triggers; pulses on the channels of modules 1 and 3 only.  CUDA == restatement (the single-module cases pin the
    restatement to the reference)."""
    from larndsim_b200 import light_sim
    p = _provider("2x2")
    z = ltu.load("2x2")
    C = ltu.consts_from_npz(z)
    C["LIGHT_TRIG_MODE"] = 0
    saved = p.light.LIGHT_TRIG_MODE
    p.light.LIGHT_TRIG_MODE = 0
    try:
        rng = np.random.default_rng(9)
        ndet, nticks = C["N_OP_CHANNEL"], 6000
        op = np.arange(ndet, dtype=np.int64)
        sig = rng.normal(0, 3, (ndet, nticks)).astype(np.float32)
        t2c = np.asarray(C["TPC_TO_OP_CHANNEL"])
        for mod, t0 in ((1, 700), (3, 1500), (3, 5200)):
            chans = t2c[C["MODULE_TO_TPCS"][mod]].ravel()
            sig[chans, t0:t0 + 100] -= (4000.0 * np.exp(-np.arange(100) / 20.0))[None, :].astype(np.float32)
        thr = ltu.thresholds(C, op)
        o_trig, o_ch, o_k = lo.get_triggers(sig, thr, op, 0, C)
        g_trig, g_ch, g_k = light_sim.get_triggers(sig, thr, op, 0)
        assert len(o_trig) >= 2 and len(set(map(tuple, o_ch))) >= 2 and np.array_equal(o_trig, g_trig) and np.array_equal(o_ch, g_ch) and np.array_equal(o_k, g_k)
        tid = np.full((ndet, nticks, 0), -1, dtype=np.int64); tph = np.zeros((ndet, nticks, 0))
        ns = 200
        d = light_sim.sim_triggers((1, 1, 1), (1, 1, 64), sig, op, tid, tph, g_trig, g_ch, ns, np.zeros((ndet, 33)))[0].cpu().numpy()
        o = lo.sim_triggers(sig, op, tid, tph, o_trig[:1], o_ch[:1, :12], ns, C)[0]
        assert np.array_equal(d[:1, :12], o) and (d != 0).sum() > 0
    finally:
        p.light.LIGHT_TRIG_MODE = saved


# ---------------------------------------------------------------------------------------- light window extent
def _extent_golden(config):
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "light_extent_%s.npz" % config))
    mode, w0, w1, tick = z["consts"]
    return z, {"LIGHT_TRIG_MODE": int(mode), "LIGHT_WINDOW": (float(w0), float(w1)), "LIGHT_TICK_SIZE": float(tick)}


def _extent_input(z, case):
    inc = ltu.extent_inputs(case)
    assert np.array_equal([inc["n_photons_det"].sum(dtype=np.float64), inc["t0_det"].sum(dtype=np.float64)], z[case + "_sum"])
    return inc


@pytest.mark.parametrize("config", ("module0", "2x2"))
def test_oracle_light_extent_matches_reference(config):
    z, C = _extent_golden(config)
    for case in ltu.EXTENT_CASES:
        inc = _extent_input(z, case)
        n, t0 = lo.get_nticks(inc, C)
        assert n == int(z[case + "_nticks"])
        assert np.asarray(t0).dtype == z[case + "_start"].dtype and np.asarray(t0) == z[case + "_start"]
        act = lo.get_active_op_channel(inc)
        assert act.dtype == np.int32 and np.array_equal(act, z[case + "_active"])


@pytest.mark.gpu
@pytest.mark.parametrize("config", ("module0", "2x2"))
def test_gpu_light_extent_identical(cuda, config):
    import torch
    from larndsim_b200 import light_sim
    z, C = _extent_golden(config)
    p = _provider(config)
    assert int(p.light.LIGHT_TRIG_MODE) == C["LIGHT_TRIG_MODE"] and tuple(p.light.LIGHT_WINDOW) == C["LIGHT_WINDOW"]
    for case in ltu.EXTENT_CASES:
        inc = _extent_input(z, case)
        n, t0 = light_sim.get_nticks(inc)
        assert n == int(z[case + "_nticks"])
        assert np.asarray(t0).dtype == z[case + "_start"].dtype and np.asarray(t0) == z[case + "_start"]
        act = light_sim.get_active_op_channel(inc)
        assert isinstance(act, torch.Tensor) and act.is_cuda and act.dtype == torch.int32
        assert np.array_equal(act.cpu().numpy(), z[case + "_active"])
    # a table large enough for many blocks: extremes planted at known places
    big = ltu.extent_inputs("lit")
    big = np.tile(big, (400, 1))
    big["t0_det"][31234, 5], big["n_photons_det"][31234, 5] = -1234.5, 1.0
    big["t0_det"][59999, 95], big["n_photons_det"][59999, 95] = 98765.25, 2.0
    big["t0_det"][100, 7], big["n_photons_det"][100, 7] = -5e6, 0.0            # no photons: must not count
    n, t0 = light_sim.get_nticks(big)
    n_ref, t0_ref = lo.get_nticks(big[[31234, 59999]], C)
    assert (n, t0) == (n_ref, t0_ref)


# ---------------------------------------------------------------------------------------- truth zero suppression
def _truth_case(z, case):
    ids, ph = ltu.truth_inputs(case)
    assert np.array_equal([float(ids.sum()), ph.sum()], z["truth_%s_sum" % case])
    return ids, ph


def _truth_modules(z, case):
    return sorted(int(k.split("_m")[-1].split("_")[0]) for k in z.files if k.startswith("truth_%s_m" % case) and k.endswith("_tick"))


def _rows_equal(rows, z, case, i_mod):
    for f in lo.TRUTH_DTYPE.names:
        want = z["truth_%s_m%d_%s" % (case, i_mod, f)]
        assert rows[f].dtype == want.dtype and np.array_equal(rows[f], want), (case, i_mod, f)


@pytest.mark.parametrize("config", ("module0", "2x2"))
def test_oracle_truth_zero_suppression_matches_reference(config):
    z, _ = _extent_golden(config)
    chan_all = z["tpc_to_op_channel"]
    for case in ltu.TRUTH_CASES:
        ids, ph = _truth_case(z, case)
        for i_mod in _truth_modules(z, case):
            chan = (chan_all[(i_mod - 1) * 2:i_mod * 2] if i_mod > 0 else chan_all).ravel()
            _rows_equal(lo.zero_suppress_waveform_truth(ids, ph, 7, 11, chan), z, case, i_mod)
    multi = z["truth_multi_m-1_trigger_id"]
    assert len(multi) > 1000 and multi[0] == 11 and multi[-1] > 2000 and (np.diff(multi) >= 0).all()      # the running sum


@pytest.mark.gpu
@pytest.mark.parametrize("config", ("module0", "2x2"))
def test_gpu_truth_zero_suppression_identical(cuda, config):
    import torch
    from larndsim_b200 import light_sim
    z, _ = _extent_golden(config)
    p = _provider(config)
    assert np.array_equal(np.asarray(p.light.TPC_TO_OP_CHANNEL), z["tpc_to_op_channel"])
    for case in ltu.TRUTH_CASES:
        ids, ph = _truth_case(z, case)
        for i_mod in _truth_modules(z, case):
            rows = light_sim.zero_suppress_waveform_truth(ids, ph, 7, 11, i_mod)
            assert isinstance(rows, np.ndarray) and rows.dtype == lo.TRUTH_DTYPE
            _rows_equal(rows, z, case, i_mod)
    # device inputs, many blocks: against the restatement's vectorised equivalent
    rng = np.random.default_rng(5)
    ids = rng.integers(0, 10**9, (4, 96, 700, 3)).astype(np.int64)
    ids[rng.random(ids.shape) < 0.7] = -1
    ph = rng.random(ids.shape)
    rows = light_sim.zero_suppress_waveform_truth(torch.from_numpy(ids).cuda(), torch.from_numpy(ph).cuda(), 3, 100)
    t, d, s, m = np.nonzero(ids != -1)
    chan = np.asarray(p.light.TPC_TO_OP_CHANNEL).ravel()
    assert len(rows) == len(t) and np.array_equal(rows["segment_id"], ids[t, d, s, m]) and np.array_equal(rows["pe_current"], ph[t, d, s, m])
    assert np.array_equal(rows["tick"], s) and np.array_equal(rows["op_channel_id"], chan[d]) and (rows["event_id"] == 3).all()
    assert np.array_equal(rows["trigger_id"], (100 + np.cumsum(t)).astype(np.int32))
