"""Drop-in for ``larndsim.active_volume`` (reference: larndsim/active_volume.py:4-46).

``select_active_volume(track_seg, tpc_borders, i_module=-1)`` returns the indices of the segments whose start or end
point lies strictly inside one of the TPC boxes.  One CUDA pass (``lsb_active_volume``) classifies every segment by the
first TPC that contains it; the index list is a prefix-sum compaction of that result.  Host records in, NumPy indices
out (as the reference); records already on the device give a torch CUDA tensor.
"""
import ctypes as C

import numpy as np
import torch

from . import _launch as _l


def _borders(tpc_borders):
    b = tpc_borders.get() if hasattr(tpc_borders, "get") else tpc_borders
    if isinstance(b, torch.Tensor):
        b = b.cpu().numpy()
    b = np.ascontiguousarray(np.asarray(b, dtype=np.float64))
    if b.ndim != 3 or b.shape[1:] != (3, 2):
        raise ValueError("tpc_borders must have shape (ntpc, 3, 2)")
    return b


def _tpc_range(n_tpc, i_module):
    if i_module < 0:
        return 0, n_tpc
    lo, hi = (i_module - 1) * 2, i_module * 2                 # 2 TPCs per module, modules count from 1
    if i_module == 0:                                         # range(-2, 0): NumPy's negative indices
        lo, hi = n_tpc - 2, n_tpc
    if lo < 0 or hi > n_tpc:
        raise IndexError("index %d is out of bounds for axis 0 with size %d" % (hi - 1, n_tpc))
    return lo, hi


def classify(track_seg, tpc_borders, tpc_lo=0, tpc_hi=None, want_indices=True):
    """(first_tpc int32[n] CUDA tensor, indices int64 CUDA tensor or None, the marshalled records)."""
    d = _l.dev(track_seg, records=True, name="track_seg")
    n = int(d.shape[0]) if len(d.shape) else 0
    b = _borders(tpc_borders)
    n_tpc = b.shape[0]
    tpc_hi = n_tpc if tpc_hi is None else tpc_hi
    bd = torch.from_numpy(b).cuda() if b.size else torch.empty(0, dtype=torch.float64, device="cuda")
    first = torch.empty(n, dtype=torch.int32, device="cuda")
    lib = _l.lib()
    lib.lsb_active_volume_ws_bytes.restype = C.c_int64
    idx = cnt = ws = None
    nws = 0
    if want_indices:
        idx = torch.empty(n, dtype=torch.int64, device="cuda")
        cnt = torch.zeros(1, dtype=torch.int64, device="cuda")
        nws = int(lib.lsb_active_volume_ws_bytes(C.c_int64(n)))
        ws = torch.empty(max(nws, 16), dtype=torch.uint8, device="cuda")
    L = _l.layout(d)
    _l.check(lib.lsb_active_volume(C.byref(L), d.c, C.c_int64(n), C.c_void_p(bd.data_ptr()), C.c_int32(n_tpc), C.c_int32(tpc_lo),
                                   C.c_int32(tpc_hi), C.c_void_p(first.data_ptr()),
                                   C.c_void_p(idx.data_ptr()) if idx is not None else None,
                                   C.c_void_p(cnt.data_ptr()) if cnt is not None else None,
                                   C.c_void_p(ws.data_ptr()) if ws is not None else None, C.c_int64(nws), _l.stream()),
              "active_volume")
    if want_indices:
        idx = idx[:int(cnt.item())]
    return first, idx, d


def select_active_volume(track_seg, tpc_borders, i_module=-1):
    b = _borders(tpc_borders)
    lo, hi = _tpc_range(b.shape[0], i_module)
    _, idx, _ = classify(track_seg, b, lo, hi, True)
    return idx.cpu().numpy() if isinstance(track_seg, np.ndarray) else idx
