"""Hit compaction + LArPix packet building on the GPU: the body of ``larndsim.fee.export_to_hdf5``
(fee.py:84-359) without the file I/O.

``export_packets(...)`` takes the arrays ``export_to_hdf5`` takes and returns ``(packets, packets_mc_ds)``:
``packets`` is a structured array (:data:`PACKET_DTYPE`, one record per LArPix packet in the reference's order:
data, timestamp, sync and trigger packets) and ``packets_mc_ds`` the ``mc_packets_assn`` table with the
reference's dtype.  The readout constants come from ``larndsim.consts`` at call time (:class:`ReadoutTables`).
The CUDA kernels are in csrc/packets.cuh, the C entry point is ``lsb_export_packets``.
"""
import ctypes as C

import numpy as np
import torch

from . import _launch as _l
from . import consts as _consts

PACKET_DTYPE = np.dtype([("packet_type", "u1"), ("io_group", "u1"), ("io_channel", "u1"), ("chip_id", "u1"),
                         ("channel_id", "u1"), ("dataword", "u1"), ("first_packet", "u1"), ("parity", "u1"),
                         ("sub_type", "u1"), ("pad", "u1", (3,)), ("receipt_timestamp", "u4"),
                         ("timestamp", "u8"), ("timestamp_s", "f8")], align=True)
assert PACKET_DTYPE.itemsize == 32
#: packet_type codes (those of larpix.format.hdf5format)
PT_DATA, PT_TIMESTAMP, PT_SYNC, PT_TRIGGER = 0, 4, 6, 7


def assn_dtype(n):
    """dtype of the ``mc_packets_assn`` dataset (fee.py:290-294)."""
    return np.dtype([("event_ids", "(1,)i8"), ("segment_ids", "(%d,)i8" % n), ("fraction", "(%d,)f8" % n),
                     ("file_traj_ids", "(%d,)i8" % n), ("fraction_traj", "(%d,)f8" % n)])


class _CTables(C.Structure):
    _fields_ = [("clock_cycle", C.c_double), ("adc_pedestal", C.c_double), ("mus", C.c_double), ("s", C.c_double),
                ("clock_reset_period", C.c_int64), ("light_trig_mode", C.c_int32),
                ("n_pixels", C.c_int32 * 2), ("n_pixels_per_tile", C.c_int32 * 2), ("n_tiles_xy", C.c_int32 * 2),
                ("n_tiles", C.c_int32), ("n_modules", C.c_int32), ("max_groups", C.c_int32), ("n_io_groups", C.c_int32),
                ("n_bad", C.c_int32),
                ("tile_map", C.c_void_p), ("tile_orientation", C.c_void_p), ("pixel_connection", C.c_void_p),
                ("tile_chip_to_io", C.c_void_p), ("module_n_groups", C.c_void_p), ("module_io_groups", C.c_void_p),
                ("io_groups", C.c_void_p), ("bad_channels", C.c_void_p)]


def _items(d):
    """(int key, value) pairs of a dict whose keys may have become strings in a JSON snapshot."""
    return [(int(k), v) for k, v in d.items()]


class ReadoutTables:
    """Flat copies of the readout constants ``export_to_hdf5`` reads: MODULE_TO_IO_GROUPS, TILE_MAP,
    TILE_ORIENTATIONS, PIXEL_CONNECTION_DICT, TILE_CHIP_TO_IO (consts/detector.py:303-356), CLOCK_CYCLE,
    CLOCK_RESET_PERIOD, LIGHT_TRIG_MODE, the unit factors and ``digitize(0)``."""

    def __init__(self, clock_cycle, clock_reset_period, light_trig_mode, adc_pedestal, mus, s, n_pixels, n_pixels_per_tile,
                 module_to_io_groups, tile_map, tile_orientations, pixel_connection, tile_chip_to_io, max_tracks_per_pixel,
                 association_count, i_mod=-1):
        self.max_tracks_per_pixel, self.association_count = int(max_tracks_per_pixel), int(association_count)
        mio = dict(_items(module_to_io_groups))
        tmap = np.ascontiguousarray(np.asarray(tile_map, dtype=np.int32))
        assert tmap.ndim == 3 and tmap.shape[0] == 2, "TILE_MAP must be [2][ntiles_x][ntiles_y]"
        tor = dict(_items(tile_orientations))
        tci = {t: dict(_items(m)) for t, m in _items(tile_chip_to_io)}
        n_tiles = max([int(tmap.max())] + list(tor) + list(tci)) + 1
        n_modules = max(mio) + 1
        max_groups = max(len(v) for v in mio.values())
        self._tile_map = tmap
        self._tile_orient = np.ones((n_tiles, 2), dtype=np.int32)
        for t, axes in tor.items():
            self._tile_orient[t] = (1 if axes[2] >= 0 else -1, 1 if axes[1] >= 0 else -1)       # rotate_tile: x_axis = axes[2], y_axis = axes[1]
        nptx, npty = int(n_pixels_per_tile[0]), int(n_pixels_per_tile[1])
        self._pix_conn = np.full((nptx, npty), -1, dtype=np.int32)
        rows = pixel_connection.items() if isinstance(pixel_connection, dict) else [((r[0], r[1]), (r[2], r[3])) for r in pixel_connection]
        for (x, y), (chip, channel) in rows:
            if 0 <= int(x) < nptx and 0 <= int(y) < npty:
                self._pix_conn[int(x), int(y)] = int(chip) * 1000 + int(channel)
        self._tile_chip_io = np.full((n_tiles, 256), -1, dtype=np.int32)
        for t, m in tci.items():
            for chip, v in m.items():
                if 0 <= chip < 256:
                    self._tile_chip_io[t, chip] = int(v)
        self._module_ng = np.zeros(n_modules, dtype=np.int32)
        self._module_io = np.zeros((n_modules, max_groups), dtype=np.int32)
        for m, groups in mio.items():
            self._module_ng[m] = len(groups)
            self._module_io[m, :len(groups)] = groups
        io_groups = np.unique(np.array([g for v in mio.values() for g in v]))                    # fee.py:120-121
        io_groups = io_groups if i_mod < 0 else io_groups[(i_mod - 1) * 2: i_mod * 2]
        self._io_groups = np.ascontiguousarray(io_groups, dtype=np.int32)
        self._bad = np.zeros(0, dtype=np.int64)
        c = _CTables()
        c.clock_cycle, c.adc_pedestal, c.mus, c.s = float(clock_cycle), float(adc_pedestal), float(mus), float(s)
        c.clock_reset_period, c.light_trig_mode = int(clock_reset_period), int(light_trig_mode)
        c.n_pixels[0], c.n_pixels[1] = int(n_pixels[0]), int(n_pixels[1])
        c.n_pixels_per_tile[0], c.n_pixels_per_tile[1] = nptx, npty
        c.n_tiles_xy[0], c.n_tiles_xy[1] = int(tmap.shape[1]), int(tmap.shape[2])
        c.n_tiles, c.n_modules, c.max_groups, c.n_io_groups = n_tiles, n_modules, max_groups, len(self._io_groups)
        self._c = c
        self._bind()

    def _bind(self):
        c = self._c
        for name, arr in (("tile_map", self._tile_map), ("tile_orientation", self._tile_orient), ("pixel_connection", self._pix_conn),
                          ("tile_chip_to_io", self._tile_chip_io), ("module_n_groups", self._module_ng),
                          ("module_io_groups", self._module_io), ("io_groups", self._io_groups), ("bad_channels", self._bad)):
            setattr(c, name, arr.ctypes.data if arr.size else None)
        c.n_bad = len(self._bad)

    def set_bad_channels(self, bad_channels):
        """``bad_channels``: {"io_group-io_channel-chip": [channel, ...]} -- the parsed YAML of fee.py:132-134 --
        or a path to that YAML file, or None."""
        if isinstance(bad_channels, str):
            import yaml
            with open(bad_channels) as f:
                bad_channels = yaml.load(f, Loader=yaml.FullLoader)
        keys = []
        for key, channels in (bad_channels or {}).items():
            g, ch, chip = (int(x) for x in str(key).split("-"))
            keys += [((g * 1000 + ch) * 1000 + chip) * 64 + int(c) for c in channels]
        self._bad = np.unique(np.asarray(keys, dtype=np.int64))
        self._bind()

    @classmethod
    def from_dict(cls, t, i_mod=-1):
        return cls(t["clock_cycle"], t["clock_reset_period"], t["light_trig_mode"], t["adc_pedestal"], t["mus"], t["s"], t["n_pixels"],
                   t["n_pixels_per_tile"], t["module_to_io_groups"], t["tile_map"], t["tile_orientations"], t["pixel_connection"],
                   t["tile_chip_to_io"], t["max_tracks_per_pixel"], t["association_count"], i_mod=i_mod)

    @classmethod
    def from_consts(cls, provider=None, i_mod=-1):
        """Read the tables from ``larndsim.consts`` (or the provider set with ``consts.use``) now."""
        p = provider or _consts.provider()
        d, li, si, un = p.detector, p.light, p.sim, p.units
        # digitize(0) (fee.py:499-515): the ADC word of an empty pixel, a constant of the configuration
        mV = un.mV
        ped = float(min(np.around(max(0.0 + d.V_PEDESTAL * mV - d.V_CM * mV, 0) * d.ADC_COUNTS / (d.V_REF * mV - d.V_CM * mV)),
                        d.ADC_COUNTS - 1))
        return cls(d.CLOCK_CYCLE, d.CLOCK_RESET_PERIOD, li.LIGHT_TRIG_MODE, ped, un.mus, un.s, d.N_PIXELS, d.N_PIXELS_PER_TILE,
                   d.MODULE_TO_IO_GROUPS, d.TILE_MAP, d.TILE_ORIENTATIONS, d.PIXEL_CONNECTION_DICT, d.TILE_CHIP_TO_IO,
                   si.MAX_TRACKS_PER_PIXEL, si.ASSOCIATION_COUNT_TO_STORE, i_mod=i_mod)


def _host(a):
    if isinstance(a, torch.Tensor):
        return a.detach().cpu().numpy()
    if isinstance(a, np.ndarray):
        return a
    if hasattr(a, "__cuda_array_interface__"):
        return torch.as_tensor(a, device="cuda").cpu().numpy()
    return np.asarray(a)


def export_packets(tables, event_id_list, adc_list, adc_ticks_list, unique_pix, current_fractions, track_ids, traj_ids,
                   event_start_times, light_trigger_times=None, light_trigger_event_id=None, light_trigger_modules=None,
                   bad_channels=None):
    """The packets and ``mc_packets_assn`` rows ``fee.export_to_hdf5`` would write, as NumPy structured arrays.
    Array arguments may be host NumPy arrays or device arrays (``__cuda_array_interface__``)."""
    if bad_channels is not None:
        tables.set_bad_channels(bad_channels)
    ev = _l.dev(event_id_list, want=np.int64, name="event_id_list")
    adc = _l.dev(adc_list, want=np.float64, name="adc_list")
    tks = _l.dev(adc_ticks_list, want=np.float64, name="adc_ticks_list")
    pix = _l.dev(unique_pix, want=np.int32, name="unique_pix")
    cf = _l.dev(current_fractions, want=np.float64, name="current_fractions")
    trk = _l.dev(track_ids, want=np.int64, name="track_ids")
    trj = _l.dev(traj_ids, want=np.int64, name="traj_ids")
    U, A = adc.shape
    K = trk.shape[1]
    if ev.shape != (U, A) or tks.shape != (U, A) or pix.shape[0] != U or cf.shape != (U, A, K) or trj.shape != (U, K):
        raise ValueError("export_packets: array shapes disagree")
    n_assn = tables.association_count
    if U == 0:
        return np.zeros(0, dtype=PACKET_DTYPE), np.zeros(0, dtype=assn_dtype(n_assn))
    # per-pixel event start (fee.py:136-137): rank of the pixel's first-slot event among the sorted unique events
    if isinstance(event_id_list, np.ndarray):
        ev0 = event_id_list[:, 0]
    else:
        ev0 = torch.as_tensor(event_id_list, device="cuda")[:, 0].cpu().numpy()          # one column, not the whole table
    _, inv = np.unique(ev0, return_inverse=True)
    est = np.asarray(_host(event_start_times), dtype=np.float64)
    t0_us = np.ascontiguousarray(est[inv])
    t0_ticks = np.ascontiguousarray((t0_us / tables._c.clock_cycle).astype(int), dtype=np.int64)
    t0u = _l.dev(t0_us, want=np.float64)
    t0t = _l.dev(t0_ticks, want=np.int64)
    n_trig = 0
    tt = te = tm = None
    if light_trigger_times is not None and len(light_trigger_times):
        tt = _l.dev(np.ascontiguousarray(_host(light_trigger_times), dtype=np.float64))
        te = _l.dev(np.ascontiguousarray(_host(light_trigger_event_id), dtype=np.int64))
        tm = _l.dev(np.ascontiguousarray(np.asarray(_host(light_trigger_modules)).astype(np.int64), dtype=np.int32))   # int(module_trig)
        n_trig = tt.shape[0]
    cap = 4 * U + 4096                 # a first guess; the call reports the exact count if it is too small
    lib = _l.lib()
    n_out = C.c_int64(0)
    adt = assn_dtype(n_assn)
    for _ in range(2):
        pk = torch.empty(cap * PACKET_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
        rows = torch.empty(cap * adt.itemsize, dtype=torch.uint8, device="cuda")          # mc_packets_assn records, written in place
        rc = lib.lsb_export_packets(C.byref(tables._c), C.c_int64(U), C.c_int32(A), C.c_int32(K), ev.c, adc.c, tks.c, pix.c, cf.c, trk.c, trj.c,
                                    t0t.c, t0u.c, C.c_int32(n_trig), tt.c if tt else None, te.c if te else None, tm.c if tm else None,
                                    C.c_int64(cap), C.c_void_p(pk.data_ptr()), C.c_void_p(rows.data_ptr()), C.c_int32(n_assn),
                                    C.byref(n_out), _l.stream())
        if rc != 0 and n_out.value > cap:
            cap = int(n_out.value)
            continue
        _l.check(rc, "export_packets")
        break
    else:
        _l.check(rc, "export_packets")
    n = int(n_out.value)
    # one D2H copy per table into pinned memory; the NumPy results are views of those buffers
    h_pk = torch.empty(n * PACKET_DTYPE.itemsize, dtype=torch.uint8, pin_memory=True)
    h_rows = torch.empty(n * adt.itemsize, dtype=torch.uint8, pin_memory=True)
    h_pk.copy_(pk[: n * PACKET_DTYPE.itemsize], non_blocking=True)
    h_rows.copy_(rows[: n * adt.itemsize], non_blocking=True)
    torch.cuda.current_stream().synchronize()
    packets = h_pk.numpy().view(PACKET_DTYPE)
    ds = h_rows.numpy().view(adt)
    return packets, ds
