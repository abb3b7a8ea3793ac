"""Golden vectors for the light-window extent (tests/golden/light_extent.npz).

Runs ONLY in the build container: imports the unmodified reference from /root/reference and calls its own
``light_sim.get_nticks`` and ``light_sim.get_active_op_channel`` (NumPy stands in for CuPy) on the seeded
light-incidence tables of tests/light_trigger_util.py, in threshold (module0) and beam (2x2) trigger mode.

    python tools/gen_golden_light_extent.py
"""
import os
import subprocess
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run(config, path):
    import refharness as rh
    rh.load_reference(simulator=False)
    consts = rh.load_properties() if config == "module0" else \
        rh.load_properties("2x2_no_modvar.yaml", "multi_tile_layout-2.4.16.yaml", "2x2_NuMI_sim_no_modvar.yaml")
    import importlib
    from larndsim import light_sim
    ls = importlib.reload(light_sim)
    import light_trigger_util as ltu
    li = consts.light
    out = {"consts": np.array([li.LIGHT_TRIG_MODE, li.LIGHT_WINDOW[0], li.LIGHT_WINDOW[1], li.LIGHT_TICK_SIZE], dtype=np.float64)}
    for case in ltu.EXTENT_CASES:
        inc = ltu.extent_inputs(case)
        n, t0 = ls.get_nticks(inc)
        out[case + "_nticks"] = np.array(n, dtype=np.int64)
        out[case + "_start"] = np.asarray(t0)                       # keeps the dtype the reference returned
        out[case + "_active"] = np.asarray(ls.get_active_op_channel(inc))
        out[case + "_sum"] = np.array([inc["n_photons_det"].sum(dtype=np.float64), inc["t0_det"].sum(dtype=np.float64)])
    for case in ltu.TRUTH_CASES:                                   # zero_suppress_waveform_truth (light_sim.py:621-661)
        ids, ph = ltu.truth_inputs(case)
        for i_mod in ((-1, 1) if config == "module0" else (-1, 3)):
            rows = ls.zero_suppress_waveform_truth(ids, ph, 7, 11, i_mod)
            for f in rows.dtype.names:
                out["truth_%s_m%d_%s" % (case, i_mod, f)] = rows[f]
        out["truth_%s_sum" % case] = np.array([ids.sum(), ph.sum()], dtype=np.float64)
    out["tpc_to_op_channel"] = np.asarray(li.TPC_TO_OP_CHANNEL)
    np.savez_compressed(path, **out)
    print(config, {k: (v.dtype.str, v.tolist() if v.size < 4 else v.shape) for k, v in out.items() if not k.startswith("truth_") or k.endswith("trigger_id")})


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(sys.argv[1], sys.argv[2])
    else:
        for config in ("module0", "2x2"):       # one fresh process each: larndsim.consts globals persist
            subprocess.check_call([sys.executable, __file__, config,
                                   os.path.join(ROOT, "tests", "golden", "light_extent_%s.npz" % config)])
