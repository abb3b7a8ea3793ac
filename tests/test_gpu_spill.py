"""The native batch loop (larndsim_b200.spill.SpillRunner -> lsb_spill_run) against the same loop written call by call with
the drop-in modules, i.e. the reference's own sequence (cli/simulate_pixels.py:667-671, 727-742, 864-1117, save_results ->
fee.export_to_hdf5): active volume cut, quench, drift, TPCBatcher masks, one chain call per (event, TPC group) batch, one
export per batch, the between-event packets in between.  Packets and mc_packets_assn rows must agree byte for byte."""
import numpy as np
import pytest

import helpers as h
from larndsim_b200 import consts as lc, synth
from larndsim_b200 import _launch as ll

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def cuda():
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch


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


@pytest.mark.parametrize("config,n,n_events,tbs", [("2x2", 6000, 3, 2), ("ndlar", 12000, 2, 2), ("module0", 3000, 2, 1)])
def test_spill_runner_matches_call_by_call_loop(cuda, config, n, n_events, tbs):
    from larndsim_b200 import spill
    mod = lc.load_snapshot(config)
    resp = synth.response_lut(mod.detector)
    tracks = synth.beam_spill_segments(n, mod.detector, seed=4242, n_events=n_events)
    tracks["segment_id"] = np.arange(len(tracks))
    tracks["file_traj_id"] = tracks["traj_id"]
    # a few segments outside every TPC: the active-volume cut must drop them
    tracks["x_start"][::97] += 1.0e4; tracks["x_end"][::97] += 1.0e4
    ref_pk, ref_rows, ref_tracks, ref_sizes = call_by_call(tracks.copy(), mod, resp, rand_seed=11, tpc_batch_size=tbs)
    for depth in (1, 3):
        runner = spill.SpillRunner(tracks.dtype, resp, depth=depth, tpc_batch_size=tbs)
        out = runner.simulate(tracks.copy(), rand_seed=11, return_tracks=True)
        assert np.array_equal(out.unit_sizes, ref_sizes)
        assert out.n_packets == len(ref_pk) and len(out.packets) == len(ref_pk)
        assert out.packets.tobytes() == ref_pk.tobytes()
        assert out.packets_mc_ds.tobytes() == ref_rows.tobytes()
        assert h.records_equal(out.tracks, ref_tracks)
        assert (out.packets["packet_type"] == 0).sum() > 100
        # the same spill again through the same runner (buffers reused): identical bytes
        again = runner.simulate(tracks.copy(), rand_seed=11)
        assert again.packets.tobytes() == ref_pk.tobytes() and again.packets_mc_ds.tobytes() == ref_rows.tobytes()
        assert out.stats["n_fma"] > 0 and out.stats["n_samples"] > 0
        runner.close()


def test_spill_runner_device_input_and_small_capacity(cuda):
    """records already on the device; a packet buffer that is too small is grown and the share re-run"""
    from larndsim_b200 import spill
    mod = lc.load_snapshot("2x2")
    resp = synth.response_lut(mod.detector)
    tracks = synth.beam_spill_segments(4000, mod.detector, seed=5, n_events=2)
    tracks["segment_id"] = np.arange(len(tracks)); tracks["file_traj_id"] = tracks["traj_id"]
    runner = spill.SpillRunner(tracks.dtype, resp, depth=2)
    a = runner.simulate(tracks.copy(), rand_seed=1)
    runner._cap = 1024                                                   # force the overflow path
    b = runner.simulate(ll.DeviceRecords(host=tracks.copy()), events=np.unique(tracks["event_id"]), rand_seed=1)
    assert len(a.packets) > 1024
    assert a.packets.tobytes() == b.packets.tobytes() and a.packets_mc_ds.tobytes() == b.packets_mc_ds.tobytes()
    c = runner.simulate(tracks.copy(), rand_seed=2)
    assert c.packets.tobytes() != a.packets.tobytes()                    # another seed, another noise realisation
    runner.close()


def test_device_rng_states_match_host(cuda):
    """create_xoroshiro128p_states on the device (GF(2) jump matrices) == the sequential host loop pinned by tests/golden/rng.npz"""
    import ctypes as C
    import torch
    from larndsim_b200 import rng
    lib = ll.lib()
    for n, seed, start in ((1, 0, 0), (1000, 1, 0), (4097, 12345678901234567, 3), (300, 2**64 - 1, 70000)):
        host = rng.create_xoroshiro128p_states_host(n, seed, start)
        dev = torch.zeros(2 * n, dtype=torch.int64, device="cuda")
        ll.check(lib.lsb_rng_create_states(C.c_void_p(dev.data_ptr()), C.c_int64(n), C.c_uint64(seed), C.c_uint64(start), ll.stream()), "rng")
        got = dev.cpu().numpy().view(np.uint64).reshape(n, 2)
        assert np.array_equal(got[:, 0], host["s0"]) and np.array_equal(got[:, 1], host["s1"])


def test_partitioned_spill_equals_single_rank(cuda):
    """2 ranks over NCCL (needs 2 GPUs): rank 0 ends up with exactly the single-rank bytes, in file order"""
    import os
    import subprocess
    import sys
    if cuda.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                        "--master-port", "29611", os.path.join(root, "tests", "spill_dist_worker.py")], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "output == single-rank output: True" in r.stdout
