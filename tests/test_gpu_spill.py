"""The native batch loop (larndsim_b200.spill.SpillRunner -> lsb_spill_run) against the same loop written call by call with
the drop-in modules, i.e. the reference's own sequence (cli/simulate_pixels.py:667-671, 727-742, 864-1117, save_results ->
fee.export_to_hdf5): active volume cut, quench, drift, TPCBatcher masks, one chain call per (event, TPC group) batch, one
export per batch, the between-event packets in between.  Packets and mc_packets_assn rows must agree byte for byte."""
import os
import sys

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


sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
from reference_loop import call_by_call  # noqa: E402  (the loop with one drop-in call per reference call)


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


def test_verbatim_batch_body_equals_fused_chain(cuda):
    """cli/simulate_pixels.py:907-1102 statement by statement with the drop-in kernels (examples/reference_loop.py
    batch_body_verbatim: dense pixels_tracks_signals, the S-iteration index-map loop, ...) == the fused chain on the same batch"""
    from reference_loop import batch_body_verbatim
    from larndsim_b200 import chain as lchain, quenching, drifting
    mod = lc.load_snapshot("2x2")
    resp = synth.response_lut(mod.detector)
    tracks = synth.beam_spill_segments(500, mod.detector, seed=77, n_events=1)
    quenching.quench[2, 256](tracks, mod.physics.BIRKS)
    drifting.drift[2, 256](tracks)
    out = batch_body_verbatim(tracks.copy(), resp, rand_seed=4, ievd=0, mod=mod)
    ch = lchain.Chain(tracks.dtype, resp, rng_fresh=True, exact_fractions=True)
    res = ch.run(ll.DeviceRecords(host=tracks.copy()), quench_mode=-1, rng_seed=4)
    assert cuda.equal(out["unique_pix"].to(cuda.int32), res.unique_pix)
    assert cuda.equal(out["track_pixel_map"], res.track_pixel_map)
    assert cuda.equal(out["pixels_signals"], res.pixels_signals)
    assert cuda.equal(out["integral_list"], res.adc_list) and cuda.equal(out["adc_tot_ticks"], res.adc_ticks_list)
    assert cuda.equal(out["adc_tot"], res.adc_digit) and cuda.equal(out["current_fractions"], res.current_fractions)
    assert int((out["adc_tot"] > 0).sum()) > 100 and not bool(out["overflow_flag"].any())
    ch.close()


def test_spill_runner_empty_and_ragged_inputs(cuda):
    """no segment at all; every segment outside the TPCs; an event whose batches are all empty; one tiny batch"""
    from larndsim_b200 import spill
    mod = lc.load_snapshot("2x2")
    resp = synth.response_lut(mod.detector)
    tracks = synth.beam_spill_segments(900, mod.detector, seed=3, n_events=3)
    tracks["segment_id"] = np.arange(len(tracks)); tracks["file_traj_id"] = tracks["traj_id"]
    runner = spill.SpillRunner(tracks.dtype, resp, depth=2)
    none = runner.simulate(tracks[:0].copy(), rand_seed=1)
    assert none.n_segments == 0 and len(none.packets) == 0 and len(none.packets_mc_ds) == 0
    far = tracks.copy()
    for f in ("x_start", "x_end", "x"):
        far[f] += 1.0e5
    out = runner.simulate(far, rand_seed=1)
    assert out.n_segments == 0 and (out.packets["packet_type"] == 0).sum() == 0
    # event 1 entirely outside: its batches are empty, the other events are unaffected
    part = tracks.copy()
    sel = part["event_id"] == 1
    for f in ("x_start", "x_end", "x"):
        part[f][sel] += 1.0e5
    a = runner.simulate(part, rand_seed=1)
    ref_pk, ref_rows, _, _ = call_by_call(part.copy(), mod, resp, rand_seed=1, tpc_batch_size=runner.tpc_batch_size)
    assert a.packets.tobytes() == ref_pk.tobytes() and a.packets_mc_ds.tobytes() == ref_rows.tobytes()
    one = runner.simulate(tracks[:3].copy(), rand_seed=1)
    assert one.n_segments == 3 and len(one.packets) >= 2
    runner.close()
