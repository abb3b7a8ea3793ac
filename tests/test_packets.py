"""Hit compaction + LArPix packet builder (reference fee.py:84-359 export_to_hdf5).

CPU: the Python restatement (oracle/packets_oracle.py) reproduces what the reference's own export_to_hdf5
produced on the committed fixtures.  GPU: the CUDA builder (through the C ABI / the fee.export_to_hdf5 drop-in)
is identical to the restatement, field for field, and to the reference fixtures."""
import numpy as np
import pytest

import packets_util as pu
from packets_util import po


@pytest.mark.parametrize("name", pu.CASES)
def test_oracle_reproduces_reference_packets(name):
    z = pu.load(name)
    packets, ds = po.export_packets(pu.tables_from_npz(z), bad_channels=pu.bad_channels_from_npz(z), **pu.inputs(z))
    assert (packets["packet_type"] == po.PT_DATA).sum() > 500
    pu.check_against_reference(packets, ds, z)


def test_oracle_rollover_case_has_rollovers():
    z = pu.load("module0_rollover_bad")
    t = pu.tables_from_npz(z)
    packets, _ = po.export_packets(t, bad_channels=pu.bad_channels_from_npz(z), **pu.inputs(z))
    # event start times beyond the 31-bit clock period: timestamps wrapped, sync + trigger packets present
    assert (z["in_event_start_times"] / t["clock_cycle"] > t["clock_reset_period"]).any()
    assert (packets["timestamp"][packets["packet_type"] != po.PT_TIMESTAMP] < t["clock_reset_period"]).all()
    assert (packets["packet_type"] == po.PT_SYNC).any() and (packets["packet_type"] == po.PT_TRIGGER).any()


def test_parity_is_odd_over_the_word():
    for chip, ch, ts, dw in ((11, 3, 12345, 87), (255, 63, 2**31 - 1, 255), (0, 0, 0, 0)):
        word = (chip << 2) | (ch << 10) | (ts << 16) | (1 << 47) | (dw << 48)
        p = po.data_parity(chip, ch, ts, 1, dw)
        assert (bin(word).count("1") + p) % 2 == 1


# ---------------------------------------------------------------------------------------- GPU
def _gpu_export(z, device_inputs=False):
    import torch
    from larndsim_b200 import packets as lp
    tables = lp.ReadoutTables.from_dict(pu.tables_from_npz(z))
    inp = pu.inputs(z)
    if device_inputs:
        for k in ("event_id_list", "adc_list", "adc_ticks_list", "unique_pix", "current_fractions", "track_ids", "traj_ids"):
            inp[k] = torch.from_numpy(np.ascontiguousarray(inp[k])).cuda()
    return lp.export_packets(tables, bad_channels=pu.bad_channels_from_npz(z), **inp)


@pytest.mark.gpu
@pytest.mark.parametrize("name", pu.CASES)
def test_gpu_packets_identical_to_oracle_and_reference(cuda, name):
    from larndsim_b200 import _launch as ll
    z = pu.load(name)
    launches0 = ll.lib().lsb_launch_count()
    packets, ds = _gpu_export(z, device_inputs=(name == "2x2"))
    assert ll.lib().lsb_launch_count() - launches0 >= 8
    o_packets, o_ds = po.export_packets(pu.tables_from_npz(z), bad_channels=pu.bad_channels_from_npz(z), **pu.inputs(z))
    assert packets.dtype == o_packets.dtype and len(packets) == len(o_packets)
    for f in packets.dtype.names:
        assert np.array_equal(packets[f], o_packets[f]), f
    for f in ds.dtype.names:
        assert np.array_equal(ds[f], o_ds[f]), f
    pu.check_against_reference(packets, ds, z)


@pytest.mark.gpu
def test_gpu_packets_edge_cases(cuda):
    from larndsim_b200 import packets as lp
    z = pu.load("module0")
    t = pu.tables_from_npz(z)
    tables = lp.ReadoutTables.from_dict(t)
    inp = pu.inputs(z)
    # no pixels at all
    empty = {k: (v[:0] if k not in ("event_start_times", "light_trigger_times", "light_trigger_event_id", "light_trigger_modules") else v)
             for k, v in inp.items()}
    packets, ds = lp.export_packets(tables, **empty)
    assert len(packets) == 0 and len(ds) == 0
    # pixels but no hit above the pedestal
    quiet = dict(inp)
    quiet["adc_list"] = np.full_like(inp["adc_list"], t["adc_pedestal"])
    packets, ds = lp.export_packets(tables, **quiet)
    assert len(packets) == 0
    # no light triggers: same as the oracle
    nolight = dict(inp, light_trigger_times=None, light_trigger_event_id=None, light_trigger_modules=None)
    packets, ds = lp.export_packets(tables, **nolight)
    o_packets, o_ds = po.export_packets(t, **nolight)
    assert len(packets) == len(o_packets) and all(np.array_equal(packets[f], o_packets[f]) for f in packets.dtype.names)
    assert all(np.array_equal(ds[f], o_ds[f]) for f in ds.dtype.names)
    # a single pixel, every ADC slot used
    one = {k: (v[:1].copy() if k not in ("event_start_times", "light_trigger_times", "light_trigger_event_id", "light_trigger_modules") else v)
           for k, v in inp.items()}
    one["adc_list"][:] = t["adc_pedestal"] + 40
    one["adc_ticks_list"][0] = np.arange(one["adc_list"].shape[1]) * 3.1
    one["event_start_times"] = inp["event_start_times"][:1]
    packets, ds = lp.export_packets(tables, **one)
    o_packets, o_ds = po.export_packets(t, **one)
    assert (packets["packet_type"] == po.PT_DATA).sum() == one["adc_list"].shape[1]
    assert all(np.array_equal(packets[f], o_packets[f]) for f in packets.dtype.names)
    assert all(np.array_equal(ds[f], o_ds[f]) for f in ds.dtype.names)


@pytest.mark.gpu
def test_gpu_packets_from_the_chain_output(cuda):
    """End of the charge chain: hits of a simulated batch -> packets; every hit above the pedestal becomes exactly one
    data packet with its ADC word, timestamps are non-decreasing within a pixel, truth rows carry the batch's segments."""
    import helpers as h
    from larndsim_b200 import chain as lchain, consts as lc, synth, _launch as ll, packets as lp
    tracks = h.production_tracks(300, "module0", 21, "cosmic")
    mod = lc.load_snapshot("module0")
    ch = lchain.Chain(tracks.dtype, synth.response_lut(mod.detector))
    res = ch.run(ll.DeviceRecords(host=tracks), rng_seed=3, n_events=1)
    z = pu.load("module0")
    t = pu.tables_from_npz(z)
    tables = lp.ReadoutTables.from_dict(t)
    U = res.n_unique_pixels
    A, K = res.adc_digit.shape[1], res.track_pixel_map.shape[1]
    ev = np.zeros((U, A), dtype=np.int64)
    traj = np.where(res.track_pixel_map.cpu().numpy() >= 0, res.track_pixel_map.cpu().numpy() // 7, -1)
    packets, ds = lp.export_packets(tables, ev, res.adc_digit, res.adc_ticks_list, res.unique_pix, res.current_fractions,
                                    res.track_pixel_map, traj, np.array([1000.0]))
    adc = res.adc_digit.cpu().numpy()
    n_hits = int((adc > t["adc_pedestal"]).sum())
    data = packets["packet_type"] == po.PT_DATA
    assert n_hits > 50 and data.sum() == n_hits
    assert np.array_equal(np.sort(packets["dataword"][data]), np.sort(adc[adc > t["adc_pedestal"]].astype(np.uint8)))
    o_packets, o_ds = po.export_packets(t, ev, adc, res.adc_ticks_list.cpu().numpy(), res.unique_pix.cpu().numpy(),
                                        res.current_fractions.cpu().numpy(), res.track_pixel_map.cpu().numpy(), traj, np.array([1000.0]))
    assert all(np.array_equal(packets[f], o_packets[f]) for f in packets.dtype.names)
    assert all(np.array_equal(ds[f], o_ds[f]) for f in ds.dtype.names)
    seg = ds["segment_ids"][data]
    assert ((seg >= -1) & (seg < len(tracks))).all() and (seg[:, 0] >= 0).all()
    ch.close()


def test_readout_tables_from_the_config_snapshot_match_the_reference_run():
    """configs/module0.json (derived from the reference YAMLs) gives the tables the reference's export used."""
    from larndsim_b200 import consts as lc, packets as lp
    z = pu.load("module0")
    a = lp.ReadoutTables.from_dict(pu.tables_from_npz(z))
    b = lp.ReadoutTables.from_consts(lc.load_snapshot("module0"))
    for name in ("_tile_map", "_tile_orient", "_pix_conn", "_tile_chip_io", "_module_ng", "_module_io", "_io_groups"):
        assert np.array_equal(getattr(a, name), getattr(b, name)), name
    for f in ("clock_cycle", "adc_pedestal", "mus", "s", "clock_reset_period", "light_trig_mode", "n_tiles", "n_modules"):
        assert getattr(a._c, f) == getattr(b._c, f), f


# ---------------------------------------------------------------------------------------- sync / timestamp / trigger
def _sync_golden(config):
    import os
    return np.load(os.path.join(pu.GOLDEN, "sync_trigger_%s.npz" % config))


def _check_small(packets, ds, z, prefix):
    kind = z[prefix + "_pk_kind"]
    assert packets.dtype == po.PACKET_DTYPE and len(packets) == len(kind) == len(ds) and len(kind) > 0
    assert np.array_equal(packets["packet_type"], kind)
    assert np.array_equal(packets["io_group"], z[prefix + "_pk_io_group"])
    ts = kind == po.PT_TIMESTAMP
    assert np.array_equal(packets["timestamp_s"][ts], z[prefix + "_pk_ts_float"][ts])            # bit-identical float64
    assert np.array_equal(packets["timestamp"][~ts].astype(np.int64), z[prefix + "_pk_timestamp"][~ts])
    assert np.array_equal(packets["sub_type"][~ts].astype(np.int64), z[prefix + "_pk_sub_type"][~ts])
    for f in ds.dtype.names:
        assert ds[f].dtype == z[prefix + "_assn_" + f].dtype or ds[f].dtype.kind == z[prefix + "_assn_" + f].dtype.kind
        assert np.array_equal(ds[f], z[prefix + "_assn_" + f])


def _small_tables(z):
    cc, period, mode, count = z["consts"]
    return dict(clock_cycle=float(cc), clock_reset_period=int(period), light_trig_mode=int(mode), association_count=int(count))


@pytest.mark.parametrize("config", ("module0", "2x2"))
def test_oracle_sync_and_trigger_packets_match_reference(config):
    from larndsim_b200 import consts as lc
    z = _sync_golden(config)
    p = lc.load_snapshot(config)
    T = _small_tables(z)
    T.update(module_to_io_groups={int(k): list(v) for k, v in dict(p.detector.MODULE_TO_IO_GROUPS).items()}, mus=p.units.mus, s=p.units.s)
    assert (p.detector.CLOCK_CYCLE, int(p.detector.CLOCK_RESET_PERIOD)) == (T["clock_cycle"], T["clock_reset_period"])
    for m in z["modules"]:
        _check_small(*po.sync_packets(T, z["in_sync"], int(m)), z, "sync%d" % m)
        _check_small(*po.timestamp_trigger_packets(T, z["in_starts"]), z, "tt%d" % m)


@pytest.mark.parametrize("config", ("module0", "2x2"))
def test_fee_sync_and_trigger_helpers_match_reference(config):
    """host helpers of the drop-in (no kernel involved): same packets, same truth rows, same warning"""
    import warnings
    from larndsim_b200 import consts as lc, fee
    z = _sync_golden(config)
    lc.load_snapshot(config)
    assert fee.get_trig_io() == int(z["trig_io"])
    for m in z["modules"]:
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            packets, ds = fee.export_sync_to_hdf5(None, z["in_sync"], int(m))
        assert len(w) == int(z["sync%d_nwarn" % m])
        _check_small(packets, ds, z, "sync%d" % m)
        _check_small(*fee.export_timestamp_trigger_to_hdf5(None, z["in_starts"], int(m)), z, "tt%d" % m)
    for t, px, py, rx, ry in z["rotate_tile"]:
        assert fee.rotate_tile((int(px), int(py)), int(t)) == (rx, ry)
    packets, ds = fee.export_sync_to_hdf5(None, np.empty(0))
    assert len(packets) == 0 and len(ds) == 0


@pytest.mark.gpu
def test_gpu_gen_event_times(cuda):
    import torch
    from larndsim_b200 import consts as lc, fee
    p = lc.load_snapshot("module0")
    t = fee.gen_event_times(20000, 5.0)
    assert t.is_cuda and t.dtype == torch.float64 and t.shape == (20000,)
    gaps = torch.diff(t).cpu().numpy()
    assert (gaps > 0).all() and float(t[0]) > 5.0
    assert abs(gaps.mean() / float(p.detector.EVENT_RATE) - 1) < 0.05            # exponential with scale EVENT_RATE
    assert abs(gaps.std() / float(p.detector.EVENT_RATE) - 1) < 0.1
