"""Kernel-by-kernel comparison of the drop-ins with the UNMODIFIED reference compiled by Numba-CUDA on the same GPU
(baseline/_ref/larndsim, installed by `pip install --no-deps --target baseline/_ref /root/reference`; test infrastructure).

Both sides are called with the reference's own launch syntax on identical inputs; constants come from the reference's
`larndsim.consts` (the drop-ins read them from there once that package is imported: the import-switch scenario of
INTEGRATION.md).  Prints one JSON object.   python tests/ref_cuda_compare.py [config] [n_segments]
"""
import json
import os
import sys
from math import ceil

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    config = sys.argv[1] if len(sys.argv) > 1 else "module0"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 300
    import ref_numba_cuda as rnc
    rnc.install_shims()
    sys.path.insert(0, rnc.REF)
    import torch
    from numba import cuda
    from numba.cuda.random import create_xoroshiro128p_states
    from larndsim import consts
    det_yaml, pix_yaml = rnc.CONFIGS[config]
    p = os.path.join(rnc.REF, "larndsim")
    consts.load_properties(os.path.join(p, "detector_properties", det_yaml), os.path.join(p, "pixel_layouts", pix_yaml),
                           os.path.join(p, "simulation_properties", "singles_sim.yaml"))
    from larndsim import quenching as rq, drifting as rd, pixels_from_track as rp, detsim as rds, fee as rf
    from larndsim.consts import detector, physics, sim
    # the drop-ins: constants now resolve to the reference's own modules (larndsim_b200.consts.provider)
    from larndsim_b200 import quenching as q, drifting as d, pixels_from_track as pf, detsim as ds, fee as f, consts as lc, synth
    import helpers as h
    assert lc.provider() is consts
    tracks = synth.cosmic_segments(n, detector, seed=31) if config == "module0" else synth.beam_spill_segments(n, detector, seed=31)
    response = synth.response_lut(detector)
    dev = torch.device("cuda", 0)
    out = {"config": config, "segments": int(n)}
    S = len(tracks)
    TPB = 256
    BPG = max(ceil(S / TPB), 1)
    # ---- quench, drift: host records in, modified in place (Numba copies back) ----
    a, b = tracks.copy(), tracks.copy()
    rq.quench[BPG, TPB](a, physics.BIRKS); rd.drift[BPG, TPB](a)
    q.quench[BPG, TPB](b, physics.BIRKS); d.drift[BPG, TPB](b)
    out["quench_drift_equal"] = h.records_equal(a, b)
    tr = b
    # ---- max_pixels, get_pixels ----
    max_radius = ceil(max(tr["tran_diff"]) * 5 / detector.PIXEL_PITCH)
    TPB = 128
    BPG = max(ceil(S / TPB), 1)
    mp_r, mp_o = np.array([0]), np.array([0])
    rp.max_pixels[BPG, TPB](tr, mp_r); pf.max_pixels[BPG, TPB](tr, mp_o)
    out["max_pixels_equal"] = bool(mp_r[0] == mp_o[0])
    P_ = int((2 * max_radius + 1) * mp_r[0] + (1 + 2 * max_radius) * max_radius * 2)
    res = {}
    for name, mod_ in (("ref", rp), ("ours", pf)):
        act = torch.full((S, int(mp_r[0])), -1, dtype=torch.int32, device=dev)
        nb = torch.full((S, P_), -1, dtype=torch.int32, device=dev)
        nr = torch.full((S, P_), -1, dtype=torch.int32, device=dev)
        npl = torch.zeros(S, dtype=torch.float64, device=dev)
        mod_.get_pixels[BPG, TPB](tr, act, nb, nr, npl, max_radius)
        torch.cuda.synchronize()
        res[name] = (act.cpu().numpy(), nb.cpu().numpy(), nr.cpu().numpy(), npl.cpu().numpy())
    out["get_pixels_equal"] = all(np.array_equal(x, y) for x, y in zip(res["ref"], res["ours"]))
    neigh = torch.from_numpy(res["ref"][1]).to(dev)
    nrad = torch.from_numpy(res["ref"][2]).to(dev)
    uniq = torch.unique(neigh.reshape(-1)); uniq = uniq[uniq != -1].contiguous()
    U = int(uniq.shape[0])
    # ---- time_intervals ----
    ti = {}
    for name, mod_ in (("ref", rds), ("ours", ds)):
        ml = torch.zeros(1, dtype=torch.int64, device=dev)
        ts = torch.empty(S, dtype=torch.float64, device=dev)
        mod_.time_intervals[BPG, TPB](ts, ml, tr)
        torch.cuda.synchronize()
        ti[name] = (ts.cpu().numpy(), int(ml.item()))
    out["time_intervals_equal"] = bool(np.array_equal(ti["ref"][0], ti["ours"][0]) and ti["ref"][1] == ti["ours"][1])
    T = ti["ref"][1]
    starts = torch.from_numpy(ti["ref"][0]).to(dev)
    # ---- tracks_current_mc with the diffusion switched off: every sample of a (segment, pixel) is deterministic, so the
    #      reference's racy state sharing between its tick threads does not matter ----
    tr0 = tr.copy(); tr0["tran_diff"] = 0; tr0["long_diff"] = 0
    TPB3 = (1, 1, 64)
    BPG3 = (S, P_, max(ceil(T / 64), 1))
    sig = {}
    d_resp = torch.from_numpy(response).to(dev)
    for name, mod_ in (("ref", rds), ("ours", ds)):
        s_ = torch.zeros((S, P_, T), dtype=torch.float32, device=dev)
        states = create_xoroshiro128p_states(S * P_, seed=3)
        mod_.tracks_current_mc[BPG3, TPB3](s_, neigh, cuda.to_device(tr0) if name == "ref" else tr0, cuda.to_device(response) if name == "ref" else d_resp, states)
        torch.cuda.synchronize()
        sig[name] = s_
    so_, sr_ = sig["ours"].cpu().numpy().astype(np.float64), sig["ref"].cpu().numpy().astype(np.float64)
    den = np.abs(sr_) + 1e-2 * np.abs(sr_).max(axis=-1, keepdims=True)
    den[den == 0] = 1.0
    pair_err = (np.abs(so_ - sr_) / den).max(axis=-1)
    valid = res["ref"][1] >= 0
    out["tracks_current_mc_sigma0_relerr"] = float(pair_err.max())
    out["tracks_current_mc_sigma0_pairs"] = int(valid.sum())
    out["tracks_current_mc_sigma0_pairs_above_1e-5"] = int((pair_err > 1e-5).sum())
    out["tracks_current_mc_sigma0_relerr_other_pairs"] = float(pair_err[pair_err <= 1e-5].max())
    out["tracks_current_mc_sigma0_support_diff_elements"] = int(((so_ != 0) != (sr_ != 0)).sum())
    out["tracks_current_mc_sigma0_charge_relerr"] = float(np.abs(so_.sum(axis=-1) - sr_.sum(axis=-1)).max() / np.abs(sr_.sum(axis=-1)).max())
    worst = np.unravel_index(np.argmax(pair_err), pair_err.shape)
    dd = np.abs(so_[worst] - sr_[worst])
    out["worst_pair"] = dict(index=[int(worst[0]), int(worst[1])], n_diff_ticks=int((dd > 1e-5 * np.abs(sr_[worst]).max()).sum()),
                             peak=float(np.abs(sr_[worst]).max()), maxdiff=float(dd.max()), first_tick=int(np.argmax(dd > 0)),
                             nonzero_ref=int((sr_[worst] != 0).sum()), nonzero_ours=int((so_[worst] != 0).sum()))
    out["tracks_current_mc_sigma0_support_equal"] = bool(torch.equal(sig["ours"] != 0, sig["ref"] != 0))
    # with diffusion: the reference's result depends on its thread interleaving; compare the collected charge per pixel row
    sgd = {}
    for name, mod_ in (("ref", rds), ("ours", ds)):
        s_ = torch.zeros((S, P_, T), dtype=torch.float32, device=dev)
        states = create_xoroshiro128p_states(S * P_, seed=3)
        mod_.tracks_current_mc[BPG3, TPB3](s_, neigh, cuda.to_device(tr) if name == "ref" else tr, cuda.to_device(response) if name == "ref" else d_resp, states)
        torch.cuda.synchronize()
        sgd[name] = s_.double().sum(dim=2).cpu().numpy()
    tot_r, tot_o = sgd["ref"].sum(), sgd["ours"].sum()
    out["tracks_current_mc_total_charge_ratio"] = float(tot_o / tot_r)
    big = np.abs(sgd["ref"]) > 0.05 * np.abs(sgd["ref"]).max()
    out["tracks_current_mc_median_pair_charge_dev"] = float(np.median(np.abs(sgd["ours"][big] / sgd["ref"][big] - 1)))
    signals = sig["ref"]
    # ---- pixel_index_map (glue), get_track_pixel_map2 ----
    pim = torch.searchsorted(uniq, neigh.clamp(min=0)).to(torch.int64)
    pim = torch.where(neigh >= 0, pim, torch.full_like(pim, -1)).contiguous()
    K = int(sim.MAX_TRACKS_PER_PIXEL)
    Upad = 32 * ceil(U / 32)                      # the reference has no index guard (detsim.py:578-580)
    uniq_p = torch.cat([uniq, torch.full((Upad - U,), -2, dtype=uniq.dtype, device=dev)])
    tpm_r = torch.full((Upad, K), -1, dtype=torch.int64, device=dev)
    rds.get_track_pixel_map2[Upad // 32, 32](tpm_r, uniq_p, neigh, nrad, int(nrad.max().item()) + 1)
    tpm_o = torch.full((U, K), -1, dtype=torch.int64, device=dev)
    ds.get_track_pixel_map2[ceil(U / 32), 32](tpm_o, uniq, neigh, nrad, int(nrad.max().item()) + 1)
    torch.cuda.synchronize()
    out["track_pixel_map2_equal"] = bool(torch.equal(tpm_r[:U], tpm_o))
    tpm = tpm_o
    # ---- sum_pixel_signals (the reference adds with float64 atomics: order varies from run to run) ----
    Tt = len(detector.TIME_TICKS)
    sums = {}
    for name, mod_ in (("ref", rds), ("ours", ds)):
        ps = torch.zeros((U, Tt), dtype=torch.float64, device=dev)
        pts = torch.zeros((U, Tt, K), dtype=torch.float64, device=dev)
        of = torch.zeros(U, dtype=torch.float64, device=dev)
        mod_.sum_pixel_signals[BPG3, TPB3](ps, signals, starts, pim, tpm, pts, of)
        torch.cuda.synchronize()
        sums[name] = (ps, pts, of)
    scale = float(sums["ref"][0].abs().max())
    out["sum_pixel_signals_maxdiff_rel"] = float((sums["ref"][0] - sums["ours"][0]).abs().max() / scale)
    out["sum_pixel_tracks_signals_maxdiff_rel"] = float((sums["ref"][1] - sums["ours"][1]).abs().max() / scale)
    out["overflow_equal"] = bool(torch.equal(sums["ref"][2], sums["ours"][2]))
    ps, pts = sums["ours"][0], sums["ours"][1]
    # ---- get_adc_values: same inputs, same RNG states (noise on), then digitize ----
    A = int(sim.MAX_ADC_VALUES)
    time_ticks = torch.linspace(0, detector.TIME_INTERVAL[1], Tt + 1, dtype=torch.float64, device=dev)
    thr = torch.full((U,), detector.DISCRIMINATION_THRESHOLD * consts.units.e, dtype=torch.float64, device=dev)
    TPB = 128
    BPG = ceil(U / TPB)
    fe = {}
    for name, mod_ in (("ref", rf), ("ours", f)):
        adc = torch.zeros((U, A), dtype=torch.float64, device=dev)
        tk = torch.zeros((U, A), dtype=torch.float64, device=dev)
        cf = torch.zeros((U, A, K), dtype=torch.float64, device=dev)
        states = create_xoroshiro128p_states(TPB * BPG, seed=9)
        mod_.get_adc_values[BPG, TPB](ps, pts, time_ticks, adc, tk, 0, states, cf, thr)
        torch.cuda.synchronize()
        fe[name] = (adc.cpu().numpy(), tk.cpu().numpy(), cf.cpu().numpy(), states.copy_to_host())
    out["adc_hits"] = int((fe["ref"][0] != 0).sum())
    out["adc_pattern_equal"] = bool(np.array_equal(fe["ref"][0] != 0, fe["ours"][0] != 0))
    out["adc_ticks_equal"] = bool(np.array_equal(fe["ref"][1], fe["ours"][1]))
    out["adc_list_equal"] = bool(np.array_equal(fe["ref"][0], fe["ours"][0]))
    out["adc_list_relerr"] = float(np.abs(fe["ref"][0] - fe["ours"][0]).max() / max(np.abs(fe["ref"][0]).max(), 1e-300))
    out["current_fractions_maxdiff"] = float(np.abs(fe["ref"][2] - fe["ours"][2]).max())
    sr, so = fe["ref"][3], fe["ours"][3]
    out["rng_states_after_equal"] = bool(np.array_equal(sr["s0"], so["s0"]) and np.array_equal(sr["s1"], so["s1"]))
    dig_o = f.digitize(fe["ours"][0])
    d_c = lc.snapshot()
    g = d_c.gain * d_c.unit_mV / d_c.unit_e
    ped = d_c.v_pedestal * d_c.unit_mV - d_c.v_cm * d_c.unit_mV
    expect = np.minimum(np.around(np.maximum(fe["ref"][0] * g + ped, 0) * d_c.adc_counts / (d_c.v_ref * d_c.unit_mV - d_c.v_cm * d_c.unit_mV)), d_c.adc_counts - 1)
    out["digitize_equal_formula"] = bool(np.array_equal(dig_o, expect))     # fee.digitize is CuPy code in the reference: its formula (fee.py:511-515)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
