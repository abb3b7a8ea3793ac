"""CPU restatement of the segment selection / batching callers -- TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may import this module.  It restates

* ``active_volume.select_active_volume`` (larndsim/active_volume.py:4-46): a segment is kept when its start OR its end
  point lies strictly inside one of the listed TPC boxes (borders sorted on the last axis first, :24; module selection
  ``i_module >= 0`` = TPCs ``2*(i_module-1) .. 2*i_module-1``, :27-28);
* ``util.batching.TPCBatcher`` (larndsim/util/batching.py:17-67): event-major iteration over groups of
  ``tpc_batch_size`` TPCs; a segment is handed out once, in the first group of its event that contains it (:49-63).

Pinned: tools/gen_golden_batching.py runs the reference's own functions (NumPy standing in for CuPy) and commits their
results under tests/golden/batching_*.npz.
"""
import numpy as np


def in_box(seg, which, box):
    """strictly inside, compared in float64 (NumPy >= 2: a float32 column against a float64 scalar promotes)"""
    x, y, z = (np.asarray(seg[a + "_" + which], dtype=np.float64) for a in "xyz")
    return (x > box[0, 0]) & (x < box[0, 1]) & (y > box[1, 0]) & (y < box[1, 1]) & (z > box[2, 0]) & (z < box[2, 1])


def select_active_volume(track_seg, tpc_borders, i_module=-1):
    borders = np.sort(np.asarray(tpc_borders, dtype=np.float64), axis=-1)
    tpcs = range(borders.shape[0]) if i_module < 0 else range((i_module - 1) * 2, i_module * 2)
    keep = np.zeros(track_seg.shape, dtype=bool)
    for t in tpcs:
        keep |= in_box(track_seg, "end", borders[t]) | in_box(track_seg, "start", borders[t])
    return np.nonzero(keep)[0]


def tpc_batches(all_track_seg, track_seg, event_separator, tpc_batch_size, tpc_borders):
    """[(event, mask), ...] in the order the reference iterator yields them"""
    borders = np.sort(np.asarray(tpc_borders, dtype=np.float64), axis=-1)
    events = np.unique(all_track_seg[event_separator])
    done = np.zeros(len(track_seg), dtype=bool)
    out = []
    for ev in events:
        for t0 in range(0, borders.shape[0], tpc_batch_size):
            inside = np.zeros(len(track_seg), dtype=bool)
            inside[select_active_volume(track_seg, borders[t0:min(t0 + tpc_batch_size, borders.shape[0])])] = True
            mask = ~done & (track_seg[event_separator] == ev) & inside
            done |= mask
            out.append((ev, mask))
    return out


def unit_of_segment(batches, n):
    """position of the batch that holds each segment (-1: none) -- the compact form the goldens store"""
    unit = np.full(n, -1, dtype=np.int32)
    for u, (_, mask) in enumerate(batches):
        assert (unit[mask] == -1).all()
        unit[mask] = u
    return unit
