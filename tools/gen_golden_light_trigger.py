"""Golden vectors for the light trigger + digitisation stage (tests/golden/light_trigger_<case>.npz).

Runs ONLY in the build container: imports the unmodified reference from /root/reference and calls its own
``light_sim.get_triggers`` and ``light_sim.sim_triggers`` (NumPy stands in for CuPy; the ``digitize_signal`` kernel
is compiled for the host from its own source with ``numba.njit`` by tools/refharness.py, i.e. with the compiled
typing of the CUDA build) on synthetic detector-response waveforms, with a zero noise spectrum (the reference draws
its noise phases from ``cupy.random``: unpinned).  The inputs are regenerated from a seed by tests/light_trigger_util.py (a checksum is stored); the constants the run
used and the outputs are stored.

    python tools/gen_golden_light_trigger.py
"""
import os
import subprocess
import sys
from math import ceil

import numpy as np

sys.path.insert(0, os.path.dirname(__file__))
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class CpArr(np.ndarray):
    """an ndarray with CuPy's ``.get()``"""
    def get(self):
        return np.asarray(self)


def run_case(name):
    import refharness as rh
    mods = rh.load_reference(simulator=False)
    if name.startswith("module0"):
        consts = rh.load_properties()
    else:
        consts = rh.load_properties("2x2_no_modvar.yaml", "multi_tile_layout-2.4.16.yaml", "2x2_NuMI_sim_no_modvar.yaml")
    import importlib
    ls = importlib.reload(mods["light_sim"])
    li, det, sim = consts.light, consts.detector, consts.sim
    host_digitize = rh.host_kernel(ls, "digitize_signal")

    class _Launch:
        def __getitem__(self, cfg):
            bpg, tpb = cfg
            return lambda *a: host_digitize(tuple(int(b) * int(t) for b, t in zip(bpg, tpb)), *a)
    ls.digitize_signal = _Launch()
    cpt = int(li.OP_CHANNEL_PER_TRIG)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import light_trigger_util as ltu
    sig, op, tid, tph = ltu.case_inputs(name, cpt, int(li.N_OP_CHANNEL))
    opc = op.astype(np.int64).view(CpArr)
    thr_all = np.repeat(np.array(li.LIGHT_TRIG_THRESHOLD)[..., np.newaxis], cpt, axis=-1).ravel()
    thr = thr_all[op].copy().reshape(-1, cpt)[..., 0]
    out = {}
    for isub in (0, 1):
        trig, trig_ch, ttype = ls.get_triggers(sig, thr, opc, isub)
        out["trig_idx_%d" % isub], out["trig_chan_%d" % isub], out["trig_type_%d" % isub] = trig, trig_ch, ttype
    trig, trig_ch = out["trig_idx_0"], out["trig_chan_0"]
    digit_samples = ceil((li.LIGHT_TRIG_WINDOW[1] + li.LIGHT_TRIG_WINDOW[0]) / li.LIGHT_DIGIT_SAMPLE_SPACING)
    TPB = (1, 1, 64)
    BPG = (max(ceil(trig.shape[0] / TPB[0]), 1), max(ceil(trig_ch.shape[1] / TPB[1]), 1), max(ceil(digit_samples / TPB[2]), 1))
    noise = np.zeros((int(li.N_OP_CHANNEL), 33))
    d, d_id, d_ph = ls.sim_triggers(BPG, TPB, sig.copy(), opc, tid, tph, trig, np.ascontiguousarray(trig_ch), digit_samples, noise)
    consts_out = dict(OP_CHANNEL_PER_TRIG=cpt, LIGHT_DIGIT_SAMPLE_SPACING=li.LIGHT_DIGIT_SAMPLE_SPACING, LIGHT_TICK_SIZE=li.LIGHT_TICK_SIZE,
                      LIGHT_TRIG_WINDOW=np.asarray(li.LIGHT_TRIG_WINDOW, dtype=np.float64), LIGHT_TRIG_MODE=int(li.LIGHT_TRIG_MODE),
                      LIGHT_NBIT=int(li.LIGHT_NBIT), MC_TRUTH_THRESHOLD=float(sim.MC_TRUTH_THRESHOLD), N_OP_CHANNEL=int(li.N_OP_CHANNEL),
                      LIGHT_TRIG_THRESHOLD=np.asarray(li.LIGHT_TRIG_THRESHOLD, dtype=np.float64),
                      OP_CHANNEL_TO_TPC=np.asarray(li.OP_CHANNEL_TO_TPC), TPC_TO_OP_CHANNEL=np.asarray(li.TPC_TO_OP_CHANNEL),
                      TPC_TO_MODULE=np.array(sorted(det.TPC_TO_MODULE.items()), dtype=np.int64),
                      MODULE_TO_TPCS=np.array([[m] + list(v) for m, v in sorted(det.MODULE_TO_TPCS.items())], dtype=np.int64))
    path = os.path.join(ROOT, "tests", "golden", "light_trigger_%s.npz" % name)
    np.savez_compressed(path, in_checksum=np.array([float(sig.astype(np.float64).sum()), float(tph.sum())]), digit_samples=digit_samples, out_digit=d, out_digit_id=d_id, out_digit_photons=d_ph,
                        **{"c_" + k: v for k, v in consts_out.items()}, **out)
    print(name, "mode", int(li.LIGHT_TRIG_MODE), "triggers", trig.tolist(), "digit", d.shape, "nonzero", int((d != 0).sum()),
          "truth entries", int((d_id >= 0).sum()), "-> %.0f kB" % (os.path.getsize(path) / 1e3))


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run_case(sys.argv[1])
    else:
        for case in ("module0", "module0_sparse", "2x2"):
            subprocess.check_call([sys.executable, os.path.abspath(__file__), case])
