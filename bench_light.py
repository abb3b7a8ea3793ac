"""Light path at the 2x2 beam-spill shape (BASELINE.json configs[3]; cli/simulate_pixels.py:745-760 and 1120-1205):
calculate_light_incidence -> sum_light_signals -> calc_scintillation_effect (16 000-tap FIR) -> calc_stat_fluctuations ->
calc_light_detector_response -> get_triggers + sim_triggers, 384 optical channels x 16 000 ticks of 1 ns, LUT smearing on,
through the drop-in modules with device arrays.  Called by bench.py (block "light") or stand-alone:

    python bench_light.py [n_segments]
"""
import ctypes as C
import json
import os
import sys
from math import ceil

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def make_inputs(n_segments=20000, seed=2024, n_prof=50):
    from larndsim_b200 import consts as lc, synth
    mod = lc.load_snapshot("2x2")
    li = mod.light
    li.ENABLE_LUT_SMEARING = True
    tracks = synth.beam_spill_segments(n_segments, mod.detector, seed=seed, n_events=1)
    tracks["segment_id"] = np.arange(len(tracks))
    lut = synth.light_lut((14, 26, 8, 48), n_prof)
    return mod, tracks, lut


def run(peak_hbm, steps=3, n_segments=20000, n_true=0):
    import torch
    from larndsim_b200 import _launch as ll, consts as lc, quenching, drifting, lightLUT, light_sim, rng
    lib = ll.lib()
    lib.lsb_profile_end.restype = C.c_int64
    mod, tracks, lut = make_inputs(n_segments)
    li = mod.light
    dev = "cuda"
    S = len(tracks)
    ndet = int(li.N_OP_CHANNEL)
    d_tr = ll.DeviceRecords(host=tracks)
    quenching.quench[1, 1](d_tr, int(mod.physics.BIRKS))
    drifting.drift[1, 1](d_tr)
    d_lut = ll.DeviceRecords(host=lut.reshape(-1))
    d_lut.shape = lut.shape
    linc_dt = np.dtype([("segment_id", "u4"), ("n_photons_det", "f4"), ("t0_det", "f4")])
    d_linc = ll.DeviceRecords(dtype=linc_dt, n=S * ndet)
    d_linc.shape = (S, ndet)
    vox = torch.zeros((S, 3), dtype=torch.int32, device=dev)
    seg_ids = torch.arange(S, dtype=torch.int64, device=dev)
    op_channel = torch.from_numpy(np.asarray(li.TPC_TO_OP_CHANNEL).ravel().astype(np.int32)).to(dev)
    nd = int(op_channel.numel())
    prof_len = float(lut["time_dist"].shape[-1])
    ev = {}

    def timed(name, fn):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record()
        torch.cuda.synchronize()
        ev.setdefault(name, []).append(e0.elapsed_time(e1))
        return r

    def once(record):
        t = timed if record else (lambda name, fn: fn())
        t("calculate_light_incidence", lambda: lightLUT.calculate_light_incidence[ceil(S / 256), 256](d_tr, d_lut, d_linc, vox))
        nticks, t_start = light_sim.get_nticks(d_linc)
        nticks = min(int(nticks), int(5e4))
        # sorted_indices: segments in descending photon yield per channel (cli/simulate_pixels.py:1144-1147; argsort glue, not timed)
        nph = torch.as_tensor(d_linc.buf.view(S * ndet, linc_dt.itemsize)[:, 4:8].contiguous().view(torch.float32).reshape(S, ndet))
        sorted_idx = torch.argsort(nph[:, op_channel.long()].t().contiguous(), dim=1, stable=True).flip(1).contiguous()
        inc = torch.zeros((nd, nticks), dtype=torch.float32, device=dev)
        tid = torch.full((nd, nticks, n_true), -1, dtype=torch.int64, device=dev)
        tph = torch.zeros((nd, nticks, n_true), dtype=torch.float64, device=dev)
        BPG, TPB = (nd, ceil(nticks / 64)), (1, 64)
        t("sum_light_signals", lambda: light_sim.sum_light_signals[BPG, TPB](d_tr, vox, seg_ids, d_linc, op_channel, d_lut, float(t_start), inc, tid, tph, sorted_idx, prof_len))
        sc, sid, sph = torch.zeros_like(inc), torch.full_like(tid, -1), torch.zeros_like(tph)
        t("calc_scintillation_effect", lambda: light_sim.calc_scintillation_effect[BPG, TPB](inc, tid, tph, sc, sid, sph))
        disc = torch.zeros_like(inc)
        states = rng.create_xoroshiro128p_states(nd * 64 * ceil(nticks / 64), 1)
        t("calc_stat_fluctuations", lambda: light_sim.calc_stat_fluctuations[BPG, TPB](sc, disc, states))
        resp, rid, rph = torch.zeros_like(inc), torch.full_like(tid, -1), torch.zeros_like(tph)
        t("calc_light_detector_response", lambda: light_sim.calc_light_detector_response[BPG, TPB](disc, sid, sph, resp, rid, rph))
        thr = np.repeat(np.asarray(li.LIGHT_TRIG_THRESHOLD)[..., np.newaxis], li.OP_CHANNEL_PER_TRIG, axis=-1).ravel()[op_channel.cpu().numpy()]
        thr = thr.reshape(-1, li.OP_CHANNEL_PER_TRIG)[..., 0]
        digit_samples = ceil((li.LIGHT_TRIG_WINDOW[1] + li.LIGHT_TRIG_WINDOW[0]) / li.LIGHT_DIGIT_SAMPLE_SPACING)

        def trig():
            tr_idx, tr_ch, _ = light_sim.get_triggers(resp, thr, op_channel, 0)
            return light_sim.sim_triggers((1, 1, 1), (1, 1, 64), resp, op_channel, rid, rph, tr_idx, tr_ch, digit_samples,
                                          np.zeros((int(li.N_OP_CHANNEL), 33))), tr_idx
        (dg, _, _), tr_idx = t("get_triggers+sim_triggers", trig)
        return dict(nticks=nticks, nonzero_inc=int((inc != 0).sum().item()), nonzero_scint=int((sc != 0).sum().item()),
                    photons_in=float(inc.double().sum().item() * li.LIGHT_TICK_SIZE), triggers=int(len(tr_idx)), digit_samples=int(dg.numel()),
                    segments_with_light=int((nph.sum(dim=1) > 0).sum().item()),
                    seg_channel_pairs=int((nph[:, op_channel.long()] > 0).sum().item()))
    once(False)
    torch.cuda.synchronize()
    lib.lsb_profile_begin(ll.stream())
    for _ in range(steps):
        shape = once(True)
    torch.cuda.synchronize()
    buf = C.create_string_buffer(1 << 16)
    lib.lsb_profile_end(buf, C.c_int64(len(buf)))
    prof = {}
    for ln in buf.value.decode().strip().split("\n"):
        f = ln.split()
        if len(f) == 3:
            prof[f[0]] = (int(f[1]), float(f[2]) / steps)
    nticks = shape["nticks"]
    taps = ceil((li.LIGHT_WINDOW[1] - li.LIGHT_WINDOW[0]) / li.LIGHT_TICK_SIZE)
    n_prof = lut["time_dist"].shape[-1]
    # sum_light_signals bytes = 4 ndet nticks + sum over (seg, det: photons > 0) of 4 n_prof + 12 S ndet;
    # FIRs: flops = 2 ndet nticks taps (skip-zero aware: taps that multiply a zero input do not count), bytes = 8 ndet nticks
    alg = {"k_sum_light_signals": ("hbm", 4.0 * nd * nticks + 4.0 * n_prof * shape["seg_channel_pairs"] + 12.0 * S * ndet),
           "k_light_fir<0>": ("flops", None), "k_light_fir<1>": ("flops", None)}
    out_k = {}
    for name, (cnt, ms) in sorted(prof.items(), key=lambda kv: -kv[1][1])[:10]:
        out_k[name] = {"launches_per_pass": cnt / steps, "ms_per_pass": ms}
    if "k_sum_light_signals" in prof:
        ms = prof["k_sum_light_signals"][1]
        b = alg["k_sum_light_signals"][1]
        out_k["k_sum_light_signals"].update({"algorithmic_bytes": b, "achieved_GBs": b / (ms * 1e-3) / 1e9, "frac_of_hbm": b / (ms * 1e-3) / 1e9 / peak_hbm})
    fir_ms = sum(v[1] for k, v in prof.items() if k.startswith("k_light_fir"))
    fir_bytes = 2 * 8.0 * nd * nticks
    result = {"workload": "2x2 config, one beam-spill event of %d segments, %d optical channels x %d ticks of %.0f ns, LUT smearing on (%d profile bins), "
                          "%d truth slots; scintillation window %d taps" % (S, nd, nticks, li.LIGHT_TICK_SIZE * 1e3, n_prof, n_true, taps),
              "ms_per_call": {k: float(np.mean(v)) for k, v in ev.items()}, "ms_total": float(sum(np.mean(v) for v in ev.values())),
              "shape": shape, "kernels": out_k,
              "fir": {"ms_both": fir_ms, "algorithmic_bytes": fir_bytes, "achieved_GBs": fir_bytes / (fir_ms * 1e-3) / 1e9 if fir_ms else None,
                      "frac_of_hbm": fir_bytes / (fir_ms * 1e-3) / 1e9 / peak_hbm if fir_ms else None,
                      "flops_skip_zero_aware": 2.0 * shape["nonzero_inc"] * taps / 2 + 2.0 * nd * nticks * 256,
                      "note": "each output tick is one float32 accumulator updated tap by tap with a rounding per add (the reference's `+=`), so the work is "
                              "a dependent chain per output: bound by FP64-add/convert latency x taps, not by HBM"}}
    return result


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
    n_true = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    print(json.dumps(run(6451.0, steps=2, n_segments=n, n_true=n_true), indent=1))
