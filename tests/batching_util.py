"""Seeded inputs for the active-volume / batching tests (shared by the golden generator and the tests)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
from larndsim_b200 import consts as lconsts, synth  # noqa: E402

# name -> (config snapshot, n segments, record kind, tpc batch sizes, seed)
CASES = {
    "module0": ("module0", 3000, "f4", (1, 2), 11),
    "2x2": ("2x2", 4000, "f4", (1, 2, 3), 12),
    "ndlar": ("ndlar", 6000, "f8", (1, 2, 4, 70), 13),
}
F8_DTYPE = np.dtype([("eventID", "i8"), ("traj_id", "i8"), ("x_start", "f8"), ("y_start", "f8"), ("z_start", "f8"),
                     ("x_end", "f8"), ("y_end", "f8"), ("z_end", "f8"), ("dE", "f8")])


def borders_of(config):
    return np.array(lconsts.load_snapshot(config).detector.TPC_BORDERS, dtype=np.float64)


def event_field(kind):
    return "event_id" if kind == "f4" else "eventID"


def segments(config, n, kind, seed):
    """Segments scattered over (and around) the detector: both ends inside one TPC, ends in different TPCs, in the gaps
    between TPCs, outside; a few coordinates sit EXACTLY on a border (the comparisons are strict); event ids are
    unsorted, non-contiguous and interleaved."""
    rng = np.random.default_rng(seed)
    b = np.sort(borders_of(config), axis=-1)
    lo, hi = b[:, :, 0].min(axis=0), b[:, :, 1].max(axis=0)
    span = hi - lo
    start = lo - 0.08 * span + rng.random((n, 3)) * 1.16 * span
    direction = rng.normal(size=(n, 3))
    direction /= np.linalg.norm(direction, axis=1, keepdims=True)
    tpc_size = (b[:, :, 1] - b[:, :, 0]).min(axis=0)
    length = rng.random(n) * np.where(rng.random(n) < 0.3, 1.5, 0.2) * tpc_size.max()
    end = start + direction * length[:, None]
    dt = synth.segment_dtype if kind == "f4" else F8_DTYPE
    seg = np.zeros(n, dtype=dt)
    for k, a in enumerate("xyz"):
        seg[a + "_start"], seg[a + "_end"] = start[:, k], end[:, k]
    # exact border hits (after the cast to the record's precision)
    for j in range(0, n, 37):
        t, a, side = rng.integers(b.shape[0]), rng.integers(3), rng.integers(2)
        which = "_start" if j % 2 else "_end"
        mid = 0.5 * (b[t, :, 0] + b[t, :, 1])
        for k, ax in enumerate("xyz"):
            seg[ax + which][j] = mid[k]
        seg["xyz"[a] + which][j] = b[t, a, side]
    ids = np.array([7, 3, 1000003, 20, 8])
    seg[event_field(kind)] = ids[rng.integers(0, len(ids), n)]
    if "traj_id" in dt.names:
        seg["traj_id"] = np.arange(n)
    return seg


def case_inputs(name):
    config, n, kind, sizes, seed = CASES[name]
    seg = segments(config, n, kind, seed)
    # the segments to batch are a subset of the file's segments: events 8 and 20 lose all / most of theirs
    ev = seg[event_field(kind)]
    keep = (ev != 8) & ~((ev == 20) & (np.arange(n) % 5 != 0))
    return seg, np.ascontiguousarray(seg[keep]), borders_of(config), sizes
