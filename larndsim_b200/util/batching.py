"""Drop-in for ``larndsim.util.batching`` (reference: larndsim/util/batching.py:1-67).

``TPCBatcher`` yields ``(event, mask)`` per (event, group of ``tpc_batch_size`` TPCs), event-major, exactly like the
reference iterator.  The reference recomputes the active-volume test over the whole segment array for every batch;
here every batch of the run is computed by ONE device pass when the iterator starts (``lsb_active_volume`` +
``lsb_batch_units``: unit key per segment, stable radix sort) and ``__next__`` only slices the result.
``units()`` gives the same batches as index arrays (what a device-resident driver wants) and ``unit_sizes`` the
segment count per batch (input of ``dist.assign_units`` for the multi-GPU partition).
"""
import ctypes as C
from math import ceil

import numpy as np
import torch

from .. import _abi, _launch as _l
from .. import active_volume as _av


class TrackSegmentBatcher(object):
    """Base class of the batchers (batching.py:6-15): holds the file's segments, the segments to batch and the name of
    the event field; subclasses are iterators over masks into ``track_seg``."""

    def __init__(self, all_track_seg, track_seg, event_separator, **kwargs):
        self.all_track_seg, self.track_seg, self.EVENT_SEPARATOR = all_track_seg, track_seg, event_separator

    def __iter__(self):
        raise NotImplementedError("subclasses define the batching scheme")


class TPCBatcher(TrackSegmentBatcher):
    def __init__(self, all_track_seg, track_seg, event_separator, tpc_batch_size=1, tpc_borders=np.empty((0, 3, 2), dtype='f4'),
                 events=None):
        """``events`` (extension): the sorted unique values of the event field, for callers whose records live on the
        device (the reference computes ``np.unique(all_track_seg[event_separator])`` on the host, batching.py:29)."""
        super().__init__(all_track_seg, track_seg, event_separator)
        self.tpc_batch_size = tpc_batch_size
        self.tpc_borders = np.sort(_av._borders(tpc_borders), axis=-1)
        self._events = np.unique(self.all_track_seg[self.EVENT_SEPARATOR]) if events is None else np.asarray(events)
        self._next_unit = 0                     # position of the iterator: unit = event rank * n_tpc_batches + TPC group
        self._order = self._offsets = None

    # -- device pass --------------------------------------------------------------------
    @property
    def n_tpc_batches(self):
        return ceil(self.tpc_borders.shape[0] / self.tpc_batch_size)

    def _plan(self):
        if self._order is not None:
            return
        first, _, d = _av.classify(self.track_seg, self.tpc_borders, want_indices=False)
        n = int(first.numel())
        nB, nE = self.n_tpc_batches, len(self._events)
        ft, off = np.dtype(d.dtype).fields[self.EVENT_SEPARATOR][:2]
        code = _abi._DTYPE_CODE.get(np.dtype(ft))
        if code is None:
            raise TypeError("field %r has unsupported dtype %s" % (self.EVENT_SEPARATOR, ft))
        ev = torch.from_numpy(np.ascontiguousarray(self._events.astype(np.int64))).cuda() if nE else \
            torch.empty(0, dtype=torch.int64, device="cuda")
        order = torch.empty(n, dtype=torch.int64, device="cuda")
        offsets = torch.zeros(nE * nB + 1, dtype=torch.int64, device="cuda")
        lib = _l.lib()
        lib.lsb_batch_units_ws_bytes.restype = C.c_int64
        nws = int(lib.lsb_batch_units_ws_bytes(C.c_int64(n)))
        ws = torch.empty(max(nws, 16), dtype=torch.uint8, device="cuda")
        _l.check(lib.lsb_batch_units(d.c, C.c_int64(n), C.c_int32(np.dtype(d.dtype).itemsize), C.c_int32(off), C.c_int32(code),
                                     C.c_void_p(ev.data_ptr()), C.c_int64(nE), C.c_void_p(first.data_ptr()),
                                     C.c_int32(self.tpc_batch_size), C.c_int32(nB), C.c_void_p(order.data_ptr()),
                                     C.c_void_p(offsets.data_ptr()), C.c_void_p(ws.data_ptr()), C.c_int64(nws), _l.stream()),
                 "batch_units")
        self._order_dev, self._offsets_dev = order, offsets
        self._offsets = offsets.cpu().numpy()
        self._order = True            # planned; the row permutation itself stays on the device until someone asks for it
        self._order_host = None
        self._n = n

    def _order_np(self):
        self._plan()
        if self._order_host is None:
            self._order_host = self._order_dev.cpu().numpy()
        return self._order_host

    @property
    def order_dev(self):
        """int64 CUDA tensor: rows of ``track_seg`` sorted by unit (stable), the rows that belong to no unit last"""
        self._plan()
        return self._order_dev

    @property
    def unit_offsets(self):
        """int64[len(self) + 1]: unit u = ``order[unit_offsets[u] : unit_offsets[u + 1]]``"""
        self._plan()
        return self._offsets

    @property
    def unit_sizes(self):
        """segments per batch, in iteration order (int64[len(self)])"""
        self._plan()
        return np.diff(self._offsets)

    def units(self, device=False):
        """(event, ascending segment indices) per batch, in iteration order"""
        self._plan()
        nB = self.n_tpc_batches
        src = self._order_dev if device else self._order_np()
        for u in range(len(self._events) * nB):
            yield self._events[u // nB], src[int(self._offsets[u]):int(self._offsets[u + 1])]

    # -- the reference's iterator protocol: (event, bool mask over track_seg), event-major, single pass ----------------
    def __len__(self):
        return len(self._events) * self.n_tpc_batches

    def __iter__(self):
        return self

    def __next__(self):
        u = self._next_unit
        if u >= len(self):
            raise StopIteration
        self._next_unit = u + 1
        self._plan()
        rows = self._order_np()[int(self._offsets[u]):int(self._offsets[u + 1])]
        mask = np.zeros(self._n, dtype=bool)
        mask[rows] = True
        return self._events[u // self.n_tpc_batches], mask
