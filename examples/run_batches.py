"""The batch loop of the reference's ``run_simulation`` (cli/simulate_pixels.py:667, 864-1117 and 1370-1390) written
with this package's drop-ins only: segments of a multi-event file -> active volume selection -> (event, TPC group)
batches -> charge chain on the GPU -> LArPix packets + ``mc_packets_assn`` rows.  Nothing here touches the reference or
the oracle; it is what a maintainer's driver looks like after switching the imports.

    python examples/run_batches.py --config 2x2 --segments 20000 --events 4
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from larndsim_b200 import consts, synth, active_volume, fee  # noqa: E402
from larndsim_b200 import chain as chain_mod, packets as packets_mod, _launch as launch, dist as dist_mod  # noqa: E402
from larndsim_b200.util import batching  # noqa: E402


def simulate(tracks, config="2x2", event_separator="event_id", tpc_batch_size=2, rand_seed=1, event_gap_us=2.0e5, chain=None):
    """-> dict(packets, packets_mc_ds, batches=[(event, n_segments, n_pixels, n_packets)], seconds, stage_seconds)"""
    import torch
    mod = consts.load_snapshot(config)
    det = mod.detector
    stage = {"setup": 0.0, "select": 0.0, "batching": 0.0, "chain": 0.0, "truth_ids": 0.0, "packets": 0.0}
    clock = [time.perf_counter()]

    def lap(name):
        now = time.perf_counter()
        stage[name] += now - clock[0]
        clock[0] = now
    # response table + device buffers: once per run (a caller that simulates several files passes its chain in)
    ch = chain if chain is not None else chain_mod.Chain(tracks.dtype, synth.response_lut(det))
    tables = packets_mod.ReadoutTables.from_consts(mod)
    torch.cuda.synchronize()
    lap("setup")
    t_start = time.perf_counter()
    # (1) keep the segments that touch an active volume                      simulate_pixels.py:667-671
    keep = active_volume.select_active_volume(tracks, det.TPC_BORDERS)
    tracks = np.ascontiguousarray(tracks[keep])
    segment_ids = tracks["segment_id"].astype(np.int64)
    trajectory_ids = tracks["traj_id"].astype(np.int64)
    events = np.unique(tracks[event_separator])
    event_times = {int(e): float(i * event_gap_us) for i, e in enumerate(events)}
    lap("select")
    # (2) every (event, TPC group) batch of the run in one device pass          simulate_pixels.py:864
    batcher = batching.TPCBatcher(tracks, tracks, event_separator, tpc_batch_size=tpc_batch_size, tpc_borders=det.TPC_BORDERS)
    sizes = batcher.unit_sizes                                                # runs the device pass
    lap("batching")
    # (2b) several GPUs: the units are independent -- longest-first assignment on the segment counts; every
    # rank computes the same plan and keeps its share.  The chain's RNG states evolve from batch to batch, so the noise
    # realisation depends on which batches a chain has seen (true of the reference's own loop order as well).
    import torch.distributed as tdist
    world = tdist.get_world_size() if tdist.is_initialized() else 1
    rank = tdist.get_rank() if tdist.is_initialized() else 0
    mine = set(dist_mod.assign_units(sizes, world)[rank]) if world > 1 else None
    nB = batcher.n_tpc_batches
    unit_ids, unit_packets, unit_rows, log = [], [], [], []
    for u, (ievd, idx) in enumerate(batcher.units()):
        if mine is not None and u not in mine:
            continue
        all_packets, all_rows = [], []
        unit_ids.append(u); unit_packets.append(all_packets); unit_rows.append(all_rows)
        if u % nB == 0:                                                        # first batch of an event: timestamp + trigger packets, :888-897
            p, r = fee.export_timestamp_trigger_to_hdf5(None, [event_times[int(ievd)]])
            all_packets.append(p); all_rows.append(r)
        if len(idx) == 0:
            log.append((int(ievd), 0, 0, 0))
            continue
        t_b = time.perf_counter()
        sub = np.ascontiguousarray(tracks[idx])
        # (3) quench -> drift -> pixels -> induced current -> pixel sums -> front end, one fused call    :918-1099
        res = ch.run(launch.DeviceRecords(host=sub), rng_seed=rand_seed + int(ievd), n_events=1)
        U = res.n_unique_pixels
        if os.environ.get("RUN_BATCHES_VERBOSE"):
            print("  batch event %d: %d segments, chain call %.2f ms" % (int(ievd), len(idx), (time.perf_counter() - t_b) * 1e3))
        lap("chain")
        if U == 0:
            log.append((int(ievd), len(idx), 0, 0))
            continue
        # (4) segment / trajectory ids of the file for the truth rows          :1110-1117
        tpm = res.track_pixel_map
        seg_of = torch.from_numpy(segment_ids[idx]).cuda()
        trj_of = torch.from_numpy(trajectory_ids[idx]).cuda()
        valid = tpm >= 0
        safe = tpm.clamp(min=0)
        track_ids = torch.where(valid, seg_of[safe], tpm)
        traj_ids = torch.where(valid, trj_of[safe], tpm)
        adc_event_ids = np.full(tuple(res.adc_digit.shape), int(ievd), dtype=np.int64)
        lap("truth_ids")
        # (5) hits -> packets                                                  :1370-1390, fee.py:84-359
        p, r = packets_mod.export_packets(tables, adc_event_ids, res.adc_digit, res.adc_ticks_list, res.unique_pix,
                                          res.current_fractions, track_ids, traj_ids, np.array([event_times[int(ievd)]]))
        all_packets.append(p); all_rows.append(r)
        log.append((int(ievd), len(idx), int(U), int(len(p))))
        lap("packets")
    if chain is None:
        ch.close()
    torch.cuda.synchronize()
    row_dtype = packets_mod.assn_dtype(int(mod.sim.ASSOCIATION_COUNT_TO_STORE))
    cat = lambda parts, dt: np.concatenate(parts) if parts else np.zeros(0, dtype=dt)      # noqa: E731
    unit_packets = [cat(p, packets_mod.PACKET_DTYPE) for p in unit_packets]
    unit_rows = [cat(r, row_dtype) for r in unit_rows]
    if world > 1:                                                             # rank 0 gets every unit back, in file order
        got_p = dist_mod.gather_unit_records(unit_ids, unit_packets, packets_mod.PACKET_DTYPE, device="cuda")
        got_r = dist_mod.gather_unit_records(unit_ids, unit_rows, row_dtype, device="cuda")
        if rank != 0:
            return dict(packets=None, packets_mc_ds=None, batches=log, seconds=time.perf_counter() - t_start, n_segments=len(tracks),
                        stage_seconds=stage, unit_sizes=sizes)
        assert got_p[0] == got_r[0] == list(range(len(sizes)))
        unit_packets, unit_rows = got_p[1], got_r[1]
    pk, rows = cat(unit_packets, packets_mod.PACKET_DTYPE), cat(unit_rows, row_dtype)
    return dict(packets=pk, packets_mc_ds=rows, batches=log, seconds=time.perf_counter() - t_start, n_segments=len(tracks),
                stage_seconds=stage, unit_sizes=sizes)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="2x2")
    ap.add_argument("--segments", type=int, default=20000)
    ap.add_argument("--events", type=int, default=4)
    ap.add_argument("--tpc-batch-size", type=int, default=2)
    a = ap.parse_args()
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1:                                                             # torchrun --nproc-per-node N examples/run_batches.py ...
        import torch.distributed as tdist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        tdist.init_process_group("nccl")
    mod = consts.load_snapshot(a.config)
    tracks = synth.beam_spill_segments(a.segments, mod.detector, seed=12345, n_events=a.events)
    tracks["segment_id"] = np.arange(len(tracks))
    ch = chain_mod.Chain(tracks.dtype, synth.response_lut(mod.detector))
    for rep in range(2):          # first pass: the chain's buffers grow to the largest batch; second pass: steady state
        out = simulate(tracks, a.config, tpc_batch_size=a.tpc_batch_size, chain=ch)
        if world > 1:
            torch.distributed.barrier()
        print("pass %d: %.3f s (rank %d of %d, %d batches here)" % (rep, out["seconds"], int(os.environ.get("RANK", "0")), world,
                                                                   len(out["batches"])), flush=True)
    ch.close()
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    if out["packets"] is None:
        return
    pk = out["packets"]
    kinds, counts = np.unique(pk["packet_type"], return_counts=True)
    print("segments %d  batches %d (non-empty %d)  packets %d %s  %.3f s after setup -> %.0f segments/s whole loop" % (
        out["n_segments"], len(out["batches"]), sum(1 for b in out["batches"] if b[1]), len(pk),
        dict(zip(kinds.tolist(), counts.tolist())), out["seconds"], out["n_segments"] / out["seconds"]))
    print("stage seconds:", {k: round(v, 4) for k, v in out["stage_seconds"].items()})


if __name__ == "__main__":
    main()
