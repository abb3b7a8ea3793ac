"""CPU restatement of the hit -> LArPix packet builder of the reference -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module; nothing here is on the
product path.  It follows ``larndsim/fee.py:84-359`` (``export_to_hdf5``) statement by statement -- the
per-pixel / per-hit loop with its clock-rollover, event-change and timestamp-packet state (:149-285) and the
``mc_packets_assn`` table (:287-342) -- but writes the packets into a structured array instead of creating
``larpix`` objects, and does no file I/O.  ``rotate_tile`` is fee.py:40-64, ``digitize(0)`` fee.py:499-515.

Pinned: tools/gen_golden_packets.py runs the reference's own ``export_to_hdf5`` (recording stand-ins for the
absent ``larpix`` / ``h5py`` packages) and commits every packet attribute it sets and the association table it
builds under tests/golden/packets_*.npz; tests/test_packets.py checks this restatement against them.
Unpinned (third party, absent here): the bit layout behind ``Packet_v2.assign_parity`` (larpix-control; the
64-bit word is restated from its documentation) and NumPy's tie order in ``np.argsort`` (ties among equal
fractions are ordered by descending slot index here, which is what a stable sort gives).
"""
import numpy as np

PACKET_DTYPE = np.dtype([("packet_type", "u1"), ("io_group", "u1"), ("io_channel", "u1"), ("chip_id", "u1"),
                         ("channel_id", "u1"), ("dataword", "u1"), ("first_packet", "u1"), ("parity", "u1"),
                         ("sub_type", "u1"), ("pad", "u1", (3,)), ("receipt_timestamp", "u4"),
                         ("timestamp", "u8"), ("timestamp_s", "f8")], align=True)
#: packet_type codes of larpix.format.hdf5format
PT_DATA, PT_TIMESTAMP, PT_SYNC, PT_TRIGGER = 0, 4, 6, 7


def assn_dtype(n):
    return np.dtype([("event_ids", "(1,)i8"), ("segment_ids", "(%d,)i8" % n), ("fraction", "(%d,)f8" % n),
                     ("file_traj_ids", "(%d,)i8" % n), ("fraction_traj", "(%d,)f8" % n)])


def data_parity(chip_id, channel_id, timestamp, first_packet, dataword):
    """Odd parity over bits 0..62 of the Packet_v2 word: type[0:2] chip[2:10] channel[10:16] timestamp[16:47]
    first_packet[47] dataword[48:56] trigger_type[56:58] fifo flags[58:62] downstream[62] (all other fields 0)."""
    word = ((chip_id & 0xFF) << 2) | ((channel_id & 0x3F) << 10) | ((timestamp & 0x7FFFFFFF) << 16) | ((first_packet & 1) << 47) | \
           ((dataword & 0xFF) << 48)
    return 1 - (bin(word).count("1") % 2)


def id2pixel(pid, n_pixels):
    """pixels_from_track.py:28-41"""
    return (pid % n_pixels[0], (pid // n_pixels[0]) % n_pixels[1], pid // (n_pixels[0] * n_pixels[1]))


def rotate_tile(pixel_id, tile_id, tables):
    axes = tables["tile_orientations"][tile_id]
    x_axis, y_axis = axes[2], axes[1]
    pix_x = pixel_id[0]
    if x_axis < 0:
        pix_x = tables["n_pixels_per_tile"][0] - pixel_id[0] - 1
    pix_y = pixel_id[1]
    if y_axis < 0:
        pix_y = tables["n_pixels_per_tile"][1] - pixel_id[1] - 1
    return pix_x, pix_y


def export_packets(tables, event_id_list, adc_list, adc_ticks_list, unique_pix, current_fractions, track_ids, traj_ids,
                   event_start_times, light_trigger_times=None, light_trigger_event_id=None, light_trigger_modules=None,
                   bad_channels=None, i_mod=-1):
    """``tables``: dict with clock_cycle, clock_reset_period, light_trig_mode, n_pixels, n_pixels_per_tile,
    module_to_io_groups {module: [io_group, ...]}, tile_map [2][nx][ny], tile_orientations {tile: (z, y, x)},
    pixel_connection {(x, y): (chip, channel)}, tile_chip_to_io {tile: {chip: io_group * 1000 + io_channel}},
    adc_pedestal (= digitize(0)), max_tracks_per_pixel, association_count, mus, s (consts.units).
    ``bad_channels``: {"io_group-io_channel-chip": [channel, ...]} (the parsed YAML of fee.py:132-134)."""
    T = tables
    CC, RESET = T["clock_cycle"], T["clock_reset_period"]
    io_groups = np.unique(np.array([g for v in T["module_to_io_groups"].values() for g in v]))
    io_groups = io_groups if i_mod < 0 else io_groups[(i_mod - 1) * 2: i_mod * 2]
    K = track_ids.shape[1]
    pk, mc_evt, mc_trk, mc_trj, mc_frac = [], [], [], [], []

    def emit(ptype, **kw):
        r = np.zeros((), dtype=PACKET_DTYPE)
        r["packet_type"] = ptype
        for k, v in kw.items():
            r[k] = v
        pk.append(r)

    def no_truth(n):
        mc_evt.append([-1]); mc_trk.append([-1] * n); mc_trj.append([-1] * n); mc_frac.append([0] * n)

    last_event = -1
    unique_events, unique_events_inv = np.unique(event_id_list[..., 0], return_inverse=True)
    event_start_time_list = (event_start_times[unique_events_inv] / CC).astype(int)
    light_trigger_times = np.empty((0,)) if light_trigger_times is None else light_trigger_times
    light_trigger_event_id = np.empty((0,), dtype=int) if light_trigger_event_id is None else light_trigger_event_id
    rollover_count = 0
    last_time_tick = -1
    for itick, adcs in enumerate(adc_list):
        ts = adc_ticks_list[itick]
        pixel_id = int(unique_pix[itick])
        pix_x, pix_y, plane_id = id2pixel(pixel_id, T["n_pixels"])
        module_id = plane_id // 2 + 1
        if module_id not in T["module_to_io_groups"]:
            continue
        tile_x = int(pix_x // T["n_pixels_per_tile"][0])
        tile_y = int(pix_y // T["n_pixels_per_tile"][1])
        anode_id = 0 if plane_id % 2 == 0 else 1
        tile_id = T["tile_map"][anode_id][tile_x][tile_y]
        for iadc, adc in enumerate(adcs):
            t = ts[iadc]
            if not adc > T["adc_pedestal"]:
                break
            while True:
                event = event_id_list[itick, iadc]
                event_t0 = event_start_time_list[itick]
                time_tick = int(np.floor(t / CC + event_t0))
                if event_t0 > RESET - 1 or time_tick > RESET - 1:
                    rollover_count += 1
                    event_start_time_list[itick:] -= RESET
                else:
                    break
            event_t0 = event_t0 % RESET
            time_tick = time_tick % RESET
            if T["light_trig_mode"] != 1:
                if event != last_event:
                    for io_group in io_groups:
                        emit(PT_TIMESTAMP, io_group=io_group, timestamp_s=event_start_times[unique_events_inv[itick]] * T["mus"] / T["s"])
                        no_truth(K)
                        emit(PT_SYNC, io_group=io_group, sub_type=ord("S"), timestamp=time_tick)
                        no_truth(K)
                    trig_mask = light_trigger_event_id == event
                    if any(trig_mask):
                        for t_trig, module_trig in zip(light_trigger_times[trig_mask], light_trigger_modules[trig_mask]):
                            t_trig = int(np.floor(t_trig / CC + event_t0)) % RESET
                            for io_group in T["module_to_io_groups"][int(module_trig)]:     # LIGHT_TRIG_MODE == 0
                                emit(PT_TRIGGER, io_group=io_group, sub_type=2, timestamp=t_trig)
                                no_truth(K)
                    last_event = event
            key = rotate_tile((pix_x % T["n_pixels_per_tile"][0], pix_y % T["n_pixels_per_tile"][1]), tile_id, T)
            if key not in T["pixel_connection"]:
                continue
            chip, channel = T["pixel_connection"][key]
            if tile_id not in T["tile_chip_to_io"] or chip not in T["tile_chip_to_io"][tile_id]:
                continue
            io_group_io_channel = T["tile_chip_to_io"][tile_id][chip]
            io_group, io_channel = io_group_io_channel // 1000, io_group_io_channel % 1000
            io_group = T["module_to_io_groups"][module_id][io_group - 1]
            chip_key = "%i-%i-%i" % (io_group, io_channel, chip)
            if bad_channels and chip_key in bad_channels and channel in bad_channels[chip_key]:
                continue
            if not time_tick == last_time_tick:
                last_time_tick = time_tick
                emit(PT_TIMESTAMP, io_group=io_group, timestamp_s=np.floor(event_start_time_list[0] * CC * T["mus"] / T["s"]))
                no_truth(T["max_tracks_per_pixel"])
            mc_evt.append([event]); mc_trk.append(track_ids[itick]); mc_trj.append(traj_ids[itick])
            mc_frac.append(current_fractions[itick][iadc])
            emit(PT_DATA, io_group=io_group, io_channel=io_channel, chip_id=chip, channel_id=channel, dataword=int(adc),
                 first_packet=1, timestamp=time_tick, receipt_timestamp=time_tick,
                 parity=data_parity(chip, channel, time_tick, 1, int(adc)))
    n = T["association_count"]
    packets = np.array(pk, dtype=PACKET_DTYPE) if pk else np.zeros(0, dtype=PACKET_DTYPE)
    ds = np.empty(len(pk), dtype=assn_dtype(n))
    if not pk:
        return packets, ds
    frac = np.array(mc_frac, dtype=np.float64)
    trk = np.array(mc_trk)
    trj = np.array(mc_trj)
    order = np.flip(np.argsort(frac, axis=1, kind="stable"), axis=1)
    a_seg = np.take_along_axis(trk, order, axis=1)
    a_trj = np.take_along_axis(trj, order, axis=1)
    a_frac = np.take_along_axis(frac, order, axis=1)

    def store(dst_ids, dst_frac, ids, fr, fill_f):
        if ids.shape[1] >= n:
            ds[dst_ids] = ids[:, :n]
            ds[dst_frac] = fr[:, :n]
        else:
            pad = n - ids.shape[1]
            ds[dst_ids] = np.pad(ids, ((0, 0), (0, pad)), mode="constant", constant_values=-1)
            ds[dst_frac] = np.pad(fr, ((0, 0), (0, pad)), mode="constant", constant_values=fill_f)
    store("segment_ids", "fraction", a_seg, a_frac, 0.0)
    t_ids = np.full(a_trj.shape, -1, dtype=np.int32)
    t_frac = np.full(a_frac.shape, 0.0, dtype=np.float32)
    for pidx, tids in enumerate(a_trj):
        mask = tids > -1
        for tidx, u in enumerate(np.unique(tids[mask])):
            t_ids[pidx][tidx] = u
            t_frac[pidx][tidx] = np.sum(a_frac[pidx][mask][tids[mask] == u])
    store("file_traj_ids", "fraction_traj", t_ids, t_frac, 0.0)
    ds["event_ids"] = np.array(mc_evt)
    return packets, ds


# ---------------------------------------------------------------------------------------------------------
# sync / timestamp + trigger packets between events: export_sync_to_hdf5 (larndsim/fee.py:361-425) and
# export_timestamp_trigger_to_hdf5 (:427-497).  Pinned by tools/gen_golden_sync_trigger.py ->
# tests/golden/sync_trigger_<config>.npz (the reference's own functions with recording stand-ins for larpix / h5py).
# ---------------------------------------------------------------------------------------------------------
def _blank_truth(n, count):
    ds = np.zeros(n, dtype=assn_dtype(count))
    ds["event_ids"] = -1
    ds["segment_ids"] = -1
    ds["file_traj_ids"] = -1
    return ds


def sync_packets(tables, sync_times, i_mod=-1):
    """one 'S' sync packet per (time, io_group); times [us] are converted to clock ticks and floored to the reset period"""
    groups = np.unique(np.array(list(tables["module_to_io_groups"].values())))
    if i_mod > 0:
        groups = tables["module_to_io_groups"][i_mod]
    out = []
    for t in np.asarray(sync_times, dtype=np.float64):
        tick = t / tables["clock_cycle"]
        period = tables["clock_reset_period"]
        if tick % period != 0:
            tick = tick // period * period
        for g in groups:
            r = np.zeros((), dtype=PACKET_DTYPE)
            r["packet_type"], r["io_group"], r["sub_type"], r["timestamp"] = PT_SYNC, g, ord("S"), int(tick)
            out.append(r)
    return (np.array(out, dtype=PACKET_DTYPE) if out else np.zeros(0, dtype=PACKET_DTYPE)), \
        _blank_truth(len(out), tables["association_count"])


def timestamp_trigger_packets(tables, event_start_times):
    """per event start time [us]: a timestamp packet [s] then a trigger packet (0x02) on the trigger io_group"""
    g = 2 if tables["light_trig_mode"] == 0 else 1
    out = []
    for t in np.asarray(event_start_times, dtype=np.float64):
        a = np.zeros((), dtype=PACKET_DTYPE)
        a["packet_type"], a["io_group"], a["timestamp_s"] = PT_TIMESTAMP, g, t * tables["mus"] / tables["s"]
        b = np.zeros((), dtype=PACKET_DTYPE)
        b["packet_type"], b["io_group"], b["sub_type"] = PT_TRIGGER, g, 2
        b["timestamp"] = int(np.floor(t / tables["clock_cycle"])) % tables["clock_reset_period"]
        out += [a, b]
    return (np.array(out, dtype=PACKET_DTYPE) if out else np.zeros(0, dtype=PACKET_DTYPE)), \
        _blank_truth(len(out), tables["association_count"])
