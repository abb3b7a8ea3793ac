"""Drop-in for ``larndsim.fee``: the device kernels (reference: larndsim/fee.py:499-655) and the hit -> LArPix
packet builder of ``export_to_hdf5`` (fee.py:84-359; GPU: packets.py / csrc/packets.cuh).  Writing the HDF5 file
itself needs the third-party ``larpix`` and ``h5py`` packages and is done only when they are importable."""
import ctypes as C

import numpy as np
import torch

from . import _launch as _l
from . import rng as _rng


def digitize(integral_list, gain=None):
    """``digitize(integral_list[, gain])`` (fee.py:499-515): integrated charge -> ADC counts.
    Accepts and returns the caller's array kind (NumPy in -> NumPy out, device in -> torch CUDA out)."""
    c = _l.snapshot()
    host = isinstance(integral_list, np.ndarray)
    q = _l.dev(integral_list, want=np.float64, name="integral_list")
    n = q.size
    g = None
    if gain is not None and not np.isscalar(gain):
        g = _l.dev(gain, want=np.float64, name="gain")
        if g.size != n:
            raise ValueError("digitize: gain must have the shape of integral_list")
    elif gain is not None:
        gt = torch.full((n,), float(gain), dtype=torch.float64, device="cuda")
        g = _l.Dev(gt.data_ptr(), (n,), np.dtype("f8"), keep=gt)
    out = torch.empty(q.shape, dtype=torch.float64, device="cuda")
    _l.check(_l.lib().lsb_digitize(C.byref(c), q.c, g.c if g is not None else None, C.c_int64(n),
                                   C.c_void_p(out.data_ptr()), _l.stream()), "digitize")
    return out.cpu().numpy() if host else out


@_l.kernel
def get_adc_values(pixels_signals, pixels_signals_tracks, time_ticks, adc_list, adc_ticks_list, time_padding, rng_states,
                   current_fractions, pixel_thresholds):
    """``get_adc_values[BPG, TPB](pixels_signals, pixels_signals_tracks, time_ticks, adc_list,
    adc_ticks_list, time_padding, rng_states, current_fractions, pixel_thresholds)`` (fee.py:517-655)."""
    c = _l.snapshot()
    ps = _l.dev(pixels_signals, want=np.float64, name="pixels_signals")
    pst = _l.dev(pixels_signals_tracks, want=np.float64, name="pixels_signals_tracks")
    tt = _l.dev(time_ticks, want=np.float64, name="time_ticks")
    adc = _l.dev(adc_list, want=np.float64, write=True, name="adc_list")
    tks = _l.dev(adc_ticks_list, want=np.float64, write=True, name="adc_ticks_list")
    st, n_rng = _rng.states_dev(rng_states)
    cf = _l.dev(current_fractions, want=np.float64, write=True, name="current_fractions")
    thr = _l.dev(pixel_thresholds, want=np.float64, name="pixel_thresholds")
    U, Tt = ps.shape
    K = cf.shape[2]
    A = adc.shape[1]
    if pst.shape != (U, Tt, K) or tks.shape != adc.shape or cf.shape[:2] != (U, A) or thr.shape[0] < U:
        raise ValueError("get_adc_values: array shapes disagree")
    _l.check(_l.lib().lsb_get_adc_values(C.byref(c), ps.c, pst.c, C.c_int64(U), C.c_int32(Tt), C.c_int32(K), tt.c,
                                         C.c_int32(tt.shape[0]), adc.c, tks.c, C.c_int32(A), C.c_double(float(time_padding)),
                                         st.c, C.c_int64(n_rng), cf.c, thr.c, _l.stream()), "get_adc_values")
    _l.finish(adc, tks, cf, st)


def export_to_hdf5(event_id_list, adc_list, adc_ticks_list, unique_pix, current_fractions, track_ids, traj_ids, filename,
                   event_start_times, light_trigger_times=None, light_trigger_event_id=None, light_trigger_modules=None,
                   bad_channels=None, i_mod=-1):
    """``export_to_hdf5(...)`` (fee.py:84-359) with the reference's arguments.  The packets (data, timestamp, sync
    and trigger packets in the reference's order, with its clock-rollover logic) and the ``mc_packets_assn`` rows
    are built on the GPU and returned as ``(packets, packets_mc_ds)``: ``packets`` is a structured array
    (``packets.PACKET_DTYPE``) instead of a list of ``larpix`` objects.  If ``filename`` is given and ``larpix`` and
    ``h5py`` are importable the file is written like the reference does (larpix ``hdf5format`` + ``mc_packets_assn``)."""
    from . import packets as _p
    tables = _p.ReadoutTables.from_consts(i_mod=i_mod)
    packets, ds = _p.export_packets(tables, event_id_list, adc_list, adc_ticks_list, unique_pix, current_fractions, track_ids,
                                    traj_ids, event_start_times, light_trigger_times, light_trigger_event_id, light_trigger_modules,
                                    bad_channels=bad_channels)
    if filename and len(packets):
        try:
            import h5py                                                  # noqa: F401
            from larpix.format import hdf5format                         # noqa: F401
        except ImportError:
            _warn_no_file(filename)
            return packets, ds
        _write_larpix_file(filename, packets, ds)
    return packets, ds


def _warn_no_file(filename):
    import warnings
    warnings.warn("larpix / h5py are not importable: %r was NOT written (the packets are returned)" % (filename,), RuntimeWarning)


def _write_larpix_file(filename, packets, ds):
    """fee.py:287-289 and :344-357 -- turn the records into larpix packet objects and append the truth table."""
    import h5py
    from larpix.packet import Packet_v2, TimestampPacket, TriggerPacket, SyncPacket, PacketCollection
    from larpix.key import Key
    from larpix.format import hdf5format
    from . import packets as _p
    from . import consts as _consts
    out = []
    for r in packets:
        t = int(r["packet_type"])
        if t == _p.PT_DATA:
            p = Packet_v2()
            p.dataword = int(r["dataword"]); p.timestamp = int(r["timestamp"])
            p.chip_key = "%i-%i-%i" % (r["io_group"], r["io_channel"], r["chip_id"])
            p.channel_id = int(r["channel_id"]); p.receipt_timestamp = int(r["receipt_timestamp"])
            p.packet_type = 0; p.first_packet = 1
            p.assign_parity()
        elif t == _p.PT_TIMESTAMP:
            p = TimestampPacket(timestamp=float(r["timestamp_s"]))
            p.chip_key = Key(int(r["io_group"]), 0, 0)
        elif t == _p.PT_SYNC:
            p = SyncPacket(sync_type=bytes([int(r["sub_type"])]), timestamp=int(r["timestamp"]), io_group=int(r["io_group"]))
        else:
            p = TriggerPacket(io_group=int(r["io_group"]), trigger_type=bytes([int(r["sub_type"])]), timestamp=int(r["timestamp"]))
        out.append(p)
    hdf5format.to_file(filename, PacketCollection(out, read_id=0, message=""), workers=1)
    d = _consts.provider().detector
    with h5py.File(filename, "a") as f:
        if "mc_packets_assn" not in f.keys():
            f.create_dataset("mc_packets_assn", data=ds, maxshape=(None,))
        else:
            f["mc_packets_assn"].resize((f["mc_packets_assn"].shape[0] + ds.shape[0]), axis=0)
            f["mc_packets_assn"][-ds.shape[0]:] = ds
        for key, val in (("vdrift", d.V_DRIFT), ("long_diff", d.LONG_DIFF), ("tran_diff", d.TRAN_DIFF),
                         ("lifetime", d.ELECTRON_LIFETIME), ("drift_length", d.DRIFT_LENGTH)):
            f["configs"].attrs[key] = val


# ---------------------------------------------------------------------------------------------------------
# the small host helpers of the reference module (a handful of records per event: no kernel involved)
# ---------------------------------------------------------------------------------------------------------
def get_trig_io():
    """``get_trig_io()`` (fee.py:28-36): io_group the light trigger is forwarded to."""
    from . import consts as _consts
    mode = _consts.provider().light.LIGHT_TRIG_MODE
    if mode == 0:
        return 2
    if mode == 1:
        return 1
    raise UnboundLocalError("cannot access local variable 'trig_io' where it is not associated with a value")


def rotate_tile(pixel_id, tile_id):
    """``rotate_tile(pixel_id, tile_id)`` (fee.py:38-62): pixel indices inside a (possibly mirrored) tile."""
    from . import consts as _consts
    d = _consts.provider().detector
    orient = d.TILE_ORIENTATIONS
    axes = orient[tile_id] if tile_id in orient else orient[str(tile_id)]
    pix_x, pix_y = pixel_id[0], pixel_id[1]
    if axes[2] < 0:
        pix_x = d.N_PIXELS_PER_TILE[0] - pixel_id[0] - 1
    if axes[1] < 0:
        pix_y = d.N_PIXELS_PER_TILE[1] - pixel_id[1] - 1
    return pix_x, pix_y


def gen_event_times(nevents, t0=None):
    """``gen_event_times(nevents, t0=NON_BEAM_EVENT_GAP)`` (fee.py:64-81): cumulative exponential gaps [us], float64 CUDA
    tensor.  The reference draws from ``cupy.random`` (unseeded, unpinned); torch's CUDA generator is used here."""
    from . import consts as _consts
    d = _consts.provider().detector
    t0 = d.NON_BEAM_EVENT_GAP if t0 is None else t0
    gaps = torch.empty(int(nevents), dtype=torch.float64, device="cuda").exponential_(1.0 / float(d.EVENT_RATE))
    return torch.cumsum(gaps, 0) + t0


def _host(a):
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    if hasattr(a, "get"):
        return np.asarray(a.get())
    return np.asarray(a)


def _no_truth_rows(n):
    from . import packets as _p
    from . import consts as _consts
    ds = np.zeros(n, dtype=_p.assn_dtype(int(_consts.provider().sim.ASSOCIATION_COUNT_TO_STORE)))
    ds["event_ids"], ds["segment_ids"], ds["file_traj_ids"] = -1, -1, -1
    return ds


def _module_io_groups(d, i_mod):
    mio = {int(k): list(v) for k, v in dict(d.MODULE_TO_IO_GROUPS).items()}
    if i_mod > 0:
        return list(mio[i_mod])
    return np.unique(np.array(list(mio.values()))).tolist()


def _finish_export(filename, packets, ds):
    if filename and len(packets):
        try:
            import h5py                                                  # noqa: F401
            from larpix.format import hdf5format                         # noqa: F401
        except ImportError:
            _warn_no_file(filename)
            return packets, ds
        _write_larpix_file(filename, packets, ds)
    return packets, ds


def export_sync_to_hdf5(filename, sync_times, i_mod=-1):
    """``export_sync_to_hdf5(filename, sync_times, i_mod=-1)`` (fee.py:361-425): one sync packet per (sync time, io_group)
    -> ``(packets, packets_mc_ds)`` as ``packets.PACKET_DTYPE`` records (the file is written when ``larpix`` / ``h5py``
    are importable).  Sync times that are not a multiple of the reset period are floored with the reference's warning.
    With no sync time the reference fails on an unbound name; empty arrays are returned here."""
    import warnings
    from . import packets as _p
    from . import consts as _consts
    d = _consts.provider().detector
    io_groups = _module_io_groups(d, i_mod)
    ticks = _host(sync_times) / d.CLOCK_CYCLE
    rows = []
    for tick in np.atleast_1d(ticks):
        if tick % d.CLOCK_RESET_PERIOD != 0:
            warnings.warn("The provided sync time is not the mutiply of the reset period!")
            tick = tick // d.CLOCK_RESET_PERIOD * d.CLOCK_RESET_PERIOD
        for g in io_groups:
            rows.append((_p.PT_SYNC, int(g), ord("S"), int(tick)))
    packets = np.zeros(len(rows), dtype=_p.PACKET_DTYPE)
    for i, (pt, g, sub, ts) in enumerate(rows):
        packets[i]["packet_type"], packets[i]["io_group"], packets[i]["sub_type"], packets[i]["timestamp"] = pt, g, sub, ts
    return _finish_export(filename, packets, _no_truth_rows(len(rows)))


def export_timestamp_trigger_to_hdf5(filename, event_start_times, i_mod=-1):
    """``export_timestamp_trigger_to_hdf5(filename, event_start_times, i_mod=-1)`` (fee.py:427-497): per event start time
    a timestamp packet [s] and a trigger packet (type 0x02, tick modulo the reset period) on the trigger io_group."""
    from . import packets as _p
    from . import consts as _consts
    p = _consts.provider()
    d, un = p.detector, p.units
    times = np.atleast_1d(_host(event_start_times))
    packets = np.zeros(2 * len(times), dtype=_p.PACKET_DTYPE)
    for i, evt_time in enumerate(times):
        g = get_trig_io()
        a, b = packets[2 * i], packets[2 * i + 1]
        a["packet_type"], a["io_group"], a["timestamp_s"] = _p.PT_TIMESTAMP, g, evt_time * un.mus / un.s
        b["packet_type"], b["io_group"], b["sub_type"] = _p.PT_TRIGGER, g, 2
        b["timestamp"] = int(np.floor(evt_time / d.CLOCK_CYCLE)) % d.CLOCK_RESET_PERIOD
    return _finish_export(filename, packets, _no_truth_rows(len(packets)))
