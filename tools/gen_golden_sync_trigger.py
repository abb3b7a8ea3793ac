"""Golden vectors for the sync / timestamp / trigger packets written between events
(tests/golden/sync_trigger_<config>.npz).

Runs ONLY in the build container: imports the unmodified reference from /root/reference and calls its own
``fee.export_sync_to_hdf5`` and ``fee.export_timestamp_trigger_to_hdf5`` with the recording stand-ins for ``larpix`` /
``h5py`` of tools/gen_golden_packets.py; also ``get_trig_io`` and ``rotate_tile`` for every tile.

    python tools/gen_golden_sync_trigger.py
"""
import os
import subprocess
import sys
import warnings

import numpy as np

sys.path.insert(0, os.path.dirname(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CONFIGS = {"module0": ("module0.yaml", "multi_tile_layout-2.3.16.yaml", "singles_sim.yaml"),
           "2x2": ("2x2_no_modvar.yaml", "multi_tile_layout-2.4.16.yaml", "2x2_NuMI_sim_no_modvar.yaml")}


def inputs(clock_cycle, reset_period):
    period_us = reset_period * clock_cycle
    sync = np.array([period_us, 2 * period_us, 3 * period_us + 12.3, 7 * period_us])        # the third one is not a multiple
    starts = np.array([0.0, 17.35, 1.2e6, period_us - 0.05, period_us + 0.05, 3.7e7 + 0.123])
    return sync, starts


def run(config, path):
    import gen_golden_packets as gp
    import refharness as rh
    gp.install_recorders()
    rh.load_reference()
    consts = rh.load_properties(*CONFIGS[config])
    from larndsim import fee
    d = consts.detector
    sync, starts = inputs(d.CLOCK_CYCLE, d.CLOCK_RESET_PERIOD)
    out = {"in_sync": sync, "in_starts": starts, "trig_io": np.array(fee.get_trig_io()),
           "consts": np.array([d.CLOCK_CYCLE, d.CLOCK_RESET_PERIOD, consts.light.LIGHT_TRIG_MODE, consts.sim.ASSOCIATION_COUNT_TO_STORE],
                              dtype=np.float64)}
    mods = [-1] + sorted(d.MODULE_TO_IO_GROUPS)
    out["modules"] = np.array(mods)
    for m in mods:
        with warnings.catch_warnings(record=True) as w:
            warnings.simplefilter("always")
            packets, ds = fee.export_sync_to_hdf5("unused.h5", sync, m)
        out["sync%d_nwarn" % m] = np.array(len(w))
        for k, v in gp.capture_to_arrays(packets).items():
            out["sync%d_pk_%s" % (m, k)] = v
        for f in ds.dtype.names:
            out["sync%d_assn_%s" % (m, f)] = ds[f]
        packets, ds = fee.export_timestamp_trigger_to_hdf5("unused.h5", starts, m)
        for k, v in gp.capture_to_arrays(packets).items():
            out["tt%d_pk_%s" % (m, k)] = v
        for f in ds.dtype.names:
            out["tt%d_assn_%s" % (m, f)] = ds[f]
    tiles = sorted(d.TILE_ORIENTATIONS)
    rot = []
    for t in tiles:
        for px, py in ((0, 0), (3, 5), (d.N_PIXELS_PER_TILE[0] - 1, d.N_PIXELS_PER_TILE[1] - 1)):
            rot.append([t, px, py, *fee.rotate_tile((px, py), t)])
    out["rotate_tile"] = np.array(rot, dtype=np.int64)
    np.savez_compressed(path, **out)
    print(config, "sync", len(out["sync-1_pk_kind"]), "tt", len(out["tt-1_pk_kind"]), "warn", int(out["sync-1_nwarn"]),
          "rot", out["rotate_tile"].shape, os.path.getsize(path))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1], sys.argv[2])
    else:
        for config in CONFIGS:                  # one fresh process each: larndsim.consts globals persist
            subprocess.check_call([sys.executable, __file__, config,
                                   os.path.join(ROOT, "tests", "golden", "sync_trigger_%s.npz" % config)])
