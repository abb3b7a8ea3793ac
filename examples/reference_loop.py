"""The reference's batch loop (cli/simulate_pixels.py:667-671, 727-742, 864-1117; save_results :179-236 ->
fee.export_to_hdf5) written call by call with the drop-in modules: one call where the reference has one call, same launch
syntax, same argument order -- what `run_simulation` executes after the import switch of INTEGRATION.md (torch stands in for
CuPy as the array library of the glue; neither the reference nor the oracle is touched).

    python examples/reference_loop.py --config 2x2 --segments 6000 --events 2

`larndsim_b200.spill.SpillRunner` runs the same loop natively (one C call per rank) and must reproduce this byte for byte:
tests/test_gpu_spill.py.
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from larndsim_b200 import _launch as ll  # noqa: E402


def call_by_call(tracks, mod, resp, rand_seed, tpc_batch_size):
    """The loop with one drop-in call per reference call."""
    import torch
    from larndsim_b200 import active_volume, quenching, drifting, chain as lchain, packets as lp, fee
    from larndsim_b200.util import batching
    det = mod.detector
    keep = active_volume.select_active_volume(tracks, det.TPC_BORDERS)
    tracks = np.ascontiguousarray(tracks[keep])
    quenching.quench[1, 1](tracks, mod.physics.BIRKS)
    drifting.drift[1, 1](tracks)
    segment_ids = tracks["segment_id"].astype(np.int64)
    trajectory_ids = tracks["file_traj_id"].astype(np.int64)
    events = np.unique(tracks["event_id"])
    event_times = (events.astype(np.int64) % mod.sim.MAX_EVENTS_PER_FILE) * float(mod.sim.SPILL_PERIOD)
    tables = lp.ReadoutTables.from_consts(mod)
    ch = lchain.Chain(tracks.dtype, resp, rng_fresh=True)
    period = det.CLOCK_RESET_PERIOD * det.CLOCK_CYCLE
    sync_start = event_times[0] // period * period + period
    packets, rows, sizes = [], [], []
    last_event = None
    nB = None
    for u, (ievd, mask) in enumerate(batching.TPCBatcher(tracks, tracks, "event_id", tpc_batch_size=tpc_batch_size, tpc_borders=det.TPC_BORDERS)):
        t0 = float(event_times[int(np.searchsorted(events, ievd))])
        if last_event is None or ievd > last_event:                            # :868-887
            if t0 - sync_start >= 0:
                sync_times = np.arange(sync_start, t0 + 1, period)
                if len(sync_times):
                    p, r = fee.export_sync_to_hdf5(None, np.full(sync_times.shape, period))
                    packets.append(p); rows.append(r)
                    sync_start = sync_times[-1] + period
            p, r = fee.export_timestamp_trigger_to_hdf5(None, [t0])
            packets.append(p); rows.append(r)
        last_event = ievd
        sub = np.ascontiguousarray(tracks[mask])
        sizes.append(len(sub))
        if len(sub) == 0:
            continue
        res = ch.run(ll.DeviceRecords(host=sub), quench_mode=-1, rng_seed=rand_seed + u, n_events=1)
        if res.n_unique_pixels == 0:
            continue
        tpm = res.track_pixel_map
        seg_of, trj_of = torch.from_numpy(segment_ids[mask]).cuda(), torch.from_numpy(trajectory_ids[mask]).cuda()
        safe = tpm.clamp(min=0)
        track_ids = torch.where(tpm >= 0, seg_of[safe], tpm)
        traj_ids = torch.where(tpm >= 0, trj_of[safe], tpm)
        ev = np.full(tuple(res.adc_digit.shape), int(ievd), dtype=np.int64)
        # charge-only run: one light trigger per event at t0, module 1 (cli/simulate_pixels.py:222-226)
        p, r = lp.export_packets(tables, ev, res.adc_digit, res.adc_ticks_list, res.unique_pix, res.current_fractions, track_ids,
                                 traj_ids, np.array([t0]), light_trigger_times=np.zeros(1), light_trigger_event_id=np.array([int(ievd)]),
                                 light_trigger_modules=np.ones(1))
        packets.append(p); rows.append(r)
    ch.close()
    return np.concatenate(packets), np.concatenate(rows), tracks, np.array(sizes)


def batch_body_verbatim(selected_tracks, response, rand_seed, ievd, mod, rng_states=None):
    """The body of the reference's batch loop, cli/simulate_pixels.py:907-1102, statement by statement with the drop-in modules
    (`cp` -> torch; the kernels are launched with the reference's own grid / block expressions).  `selected_tracks` is the
    host record array of one batch, already quenched and drifted.  Returns the arrays the loop appends to `results_acc`."""
    import torch
    from math import ceil
    from larndsim_b200 import pixels_from_track, detsim, fee
    from larndsim_b200.rng import maybe_create_rng_states
    detector, sim, consts_units = mod.detector, mod.sim, mod.units
    dev = "cuda"
    itrk = 0
    event_ids = selected_tracks[sim.EVENT_SEPARATOR]
    unique_eventIDs = np.unique(event_ids)
    # max_pixels (:917-928)
    max_radius = ceil(max(selected_tracks["tran_diff"]) * 5 / detector.PIXEL_PITCH)
    TPB = 128
    BPG = max(ceil(selected_tracks.shape[0] / TPB), 1)
    max_pixels = np.array([0])
    pixels_from_track.max_pixels[BPG, TPB](selected_tracks, max_pixels)
    max_neighboring_pixels = (2 * max_radius + 1) * max_pixels[0] + (1 + 2 * max_radius) * max_radius * 2
    active_pixels = torch.full((selected_tracks.shape[0], int(max_pixels[0])), -1, dtype=torch.int32, device=dev)
    neighboring_pixels = torch.full((selected_tracks.shape[0], int(max_neighboring_pixels)), -1, dtype=torch.int32, device=dev)
    neighboring_radius = torch.full((selected_tracks.shape[0], int(max_neighboring_pixels)), -1, dtype=torch.int32, device=dev)
    n_pixels_list = torch.zeros(selected_tracks.shape[0], dtype=torch.float64, device=dev)
    # get_pixels (:943-950)
    pixels_from_track.get_pixels[BPG, TPB](selected_tracks, active_pixels, neighboring_pixels, neighboring_radius, n_pixels_list, max_radius)
    # unique_pix (:952-956)
    shapes = neighboring_pixels.shape
    joined = neighboring_pixels.reshape(shapes[0] * shapes[1])
    unique_pix = torch.unique(joined)
    unique_pix = unique_pix[(unique_pix != -1)]
    # time_intervals (:997-1002)
    max_length = torch.tensor([0], device=dev)
    track_starts = torch.empty(selected_tracks.shape[0], dtype=torch.float64, device=dev)
    detsim.time_intervals[BPG, TPB](track_starts, max_length, selected_tracks)
    # tracks_current (:1004-1016)
    signals = torch.zeros((selected_tracks.shape[0], neighboring_pixels.shape[1], int(max_length.cpu().numpy()[0])), dtype=torch.float32, device=dev)
    TPB = (1, 1, 64)
    BPG_X = max(ceil(signals.shape[0] / TPB[0]), 1)
    BPG_Y = max(ceil(signals.shape[1] / TPB[1]), 1)
    BPG_Z = max(ceil(signals.shape[2] / TPB[2]), 1)
    BPG = (BPG_X, BPG_Y, BPG_Z)
    rng_states = maybe_create_rng_states(int(np.prod(TPB[:2]) * np.prod(BPG[:2])), seed=rand_seed + ievd + itrk, rng_states=rng_states)
    detsim.tracks_current_mc[BPG, TPB](signals, neighboring_pixels, selected_tracks, response, rng_states)
    # pixel_index_map (:1019-1025)
    pixel_index_map = torch.full((selected_tracks.shape[0], neighboring_pixels.shape[1]), -1, dtype=torch.int64, device=dev)
    for i_ in range(selected_tracks.shape[0]):
        compare = neighboring_pixels[i_, ..., None] == unique_pix
        indices = torch.where(compare)
        pixel_index_map[i_, indices[0]] = indices[1]
    # track_pixel_map (:1028-1042)
    max_segments_to_trace = sim.MAX_TRACKS_PER_PIXEL
    track_pixel_map = torch.full((unique_pix.shape[0], max_segments_to_trace), -1, dtype=torch.int64, device=dev)
    TPB = 32
    BPG = max(ceil(unique_pix.shape[0] / TPB), 1)
    detsim.get_track_pixel_map2[BPG, TPB](track_pixel_map, unique_pix, neighboring_pixels, neighboring_radius, neighboring_radius.max().item() + 1)
    # sum_pixels_signals (:1045-1067)
    TPB = (1, 1, 64)
    BPG = (BPG_X, BPG_Y, BPG_Z)
    pixels_signals = torch.zeros((len(unique_pix), len(detector.TIME_TICKS)), dtype=torch.float64, device=dev)
    pixels_tracks_signals = torch.zeros((len(unique_pix), len(detector.TIME_TICKS), track_pixel_map.shape[1]), dtype=torch.float64, device=dev)
    overflow_flag = torch.zeros(len(unique_pix), dtype=torch.float64, device=dev)
    detsim.sum_pixel_signals[BPG, TPB](pixels_signals, signals, track_starts, pixel_index_map, track_pixel_map, pixels_tracks_signals, overflow_flag)
    # get_adc_values (:1070-1102)
    # cp.linspace follows numpy's formula (start + i * step, last element = stop); torch.linspace does not
    time_ticks = torch.from_numpy(np.linspace(0, len(unique_eventIDs) * detector.TIME_INTERVAL[1], pixels_signals.shape[1] + 1)).to(dev)
    integral_list = torch.zeros((pixels_signals.shape[0], sim.MAX_ADC_VALUES), dtype=torch.float64, device=dev)
    adc_ticks_list = torch.zeros((pixels_signals.shape[0], sim.MAX_ADC_VALUES), dtype=torch.float64, device=dev)
    current_fractions = torch.zeros((pixels_signals.shape[0], sim.MAX_ADC_VALUES, track_pixel_map.shape[1]), dtype=torch.float64, device=dev)
    TPB = 128
    BPG = ceil(pixels_signals.shape[0] / TPB)
    rng_states = maybe_create_rng_states(int(TPB * BPG), seed=rand_seed + ievd + itrk, rng_states=rng_states)
    pixel_thresholds = torch.full((pixels_signals.shape[0],), detector.DISCRIMINATION_THRESHOLD * consts_units.e, dtype=torch.float64, device=dev)
    fee.get_adc_values[BPG, TPB](pixels_signals, pixels_tracks_signals, time_ticks, integral_list, adc_ticks_list, 0, rng_states, current_fractions, pixel_thresholds)
    adc_list = fee.digitize(integral_list)
    return dict(unique_pix=unique_pix, adc_tot=adc_list, adc_tot_ticks=adc_ticks_list, current_fractions=current_fractions,
                track_pixel_map=track_pixel_map, integral_list=integral_list, pixels_signals=pixels_signals, overflow_flag=overflow_flag,
                rng_states=rng_states)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="2x2")
    ap.add_argument("--segments", type=int, default=6000)
    ap.add_argument("--events", type=int, default=2)
    ap.add_argument("--tpc-batch-size", type=int, default=2)
    a = ap.parse_args()
    from larndsim_b200 import consts as lc, synth
    mod = lc.load_snapshot(a.config)
    tracks = synth.beam_spill_segments(a.segments, mod.detector, seed=12345, n_events=a.events)
    tracks["segment_id"] = np.arange(len(tracks))
    tracks["file_traj_id"] = tracks["traj_id"]
    pk, rows, tr, sizes = call_by_call(tracks, mod, synth.response_lut(mod.detector), rand_seed=1, tpc_batch_size=a.tpc_batch_size)
    kinds, counts = np.unique(pk["packet_type"], return_counts=True)
    print("%d segments kept, %d batches (%d non-empty), %d packets %s, %d truth rows" % (
        len(tr), len(sizes), int((sizes > 0).sum()), len(pk), dict(zip(kinds.tolist(), counts.tolist())), len(rows)))


if __name__ == "__main__":
    main()
