"""Golden vectors for segment selection + batching (tests/golden/batching_<case>.npz).

Runs ONLY in the build container: imports the unmodified reference from /root/reference and calls its own
``active_volume.select_active_volume`` and ``util.batching.TPCBatcher`` (NumPy stands in for CuPy) on the seeded
segments of tests/batching_util.py.  Stored: a checksum of the inputs, the selected indices (all TPCs and per module),
and for every TPC batch size the yielded events and, per segment, the position of the batch that returned it.

    python tools/gen_golden_batching.py
"""
import hashlib
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import refharness as rh
    rh.load_reference(simulator=False)
    from larndsim import active_volume
    from larndsim.util import batching
    import batching_util as bu
    for name in bu.CASES:
        all_seg, seg, borders, sizes = bu.case_inputs(name)
        kind = bu.CASES[name][2]
        sep = bu.event_field(kind)
        out = {"checksum": np.frombuffer(hashlib.sha256(all_seg.tobytes() + borders.tobytes()).digest(), dtype=np.uint8)}
        out["sel_all"] = active_volume.select_active_volume(all_seg, borders)
        n_mod = borders.shape[0] // 2
        for m in sorted({1, n_mod, max(1, n_mod // 2)}):
            out["sel_module_%d" % m] = active_volume.select_active_volume(all_seg, borders, m)
        for bs in sizes:
            it = batching.TPCBatcher(all_seg, seg, sep, tpc_batch_size=bs, tpc_borders=borders)
            events, unit = [], np.full(len(seg), -1, dtype=np.int32)
            assert len(it) == len(np.unique(all_seg[sep])) * -(-borders.shape[0] // bs)
            for u, (ev, mask) in enumerate(it):
                events.append(ev)
                assert mask.dtype == bool and mask.shape == seg.shape and (unit[mask] == -1).all()
                unit[mask] = u
            out["events_bs%d" % bs] = np.array(events)
            out["unit_bs%d" % bs] = unit
        path = os.path.join(ROOT, "tests", "golden", "batching_%s.npz" % name)
        np.savez_compressed(path, **out)
        print(name, {k: (v.shape, int((v >= 0).sum()) if k.startswith("unit") else "") for k, v in out.items()}, os.path.getsize(path))


if __name__ == "__main__":
    main()
