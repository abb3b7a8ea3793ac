"""Inputs of the light trigger / digitisation tests.  The synthetic detector-response waveforms are regenerated
from a seed (they are noise and would not compress); tests/golden/light_trigger_<case>.npz holds the constants of
the reference run and everything the reference's get_triggers / sim_triggers returned for them
(tools/gen_golden_light_trigger.py)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

CASES = ("module0", "module0_sparse", "2x2")
SEEDS = {"module0": 101, "module0_sparse": 202, "2x2": 303}


def waveforms(rng, ndet, nticks, pulses, M, f8=False):
    sig = rng.normal(0, 4, (ndet, nticks)).astype(np.float64 if f8 else np.float32)
    tid = np.full((ndet, nticks, M), -1, dtype=np.int64)
    tph = np.zeros((ndet, nticks, M), dtype=np.float64)
    for k, (t0, amp, chans) in enumerate(pulses):
        n = min(120, nticks - t0)
        shape = np.exp(-np.arange(n) / 25.0)
        sig[chans, t0:t0 + n] -= (amp * shape)[None, :].astype(sig.dtype)
        if M:
            tid[chans, t0:t0 + n, 0] = 100 + k
            tph[chans, t0:t0 + n, 0] = rng.uniform(0, 3, (len(chans), n))
            if M > 1:
                half = chans[::2]
                tid[half, t0 + 5:t0 + n - 5, 1] = 200 + k
                tph[half, t0 + 5:t0 + n - 5, 1] = rng.uniform(0, 0.3, (len(half), n - 10))
    return sig, tid, tph


def case_inputs(name, cpt, n_op_channel):
    """(signal, op_channel, truth ids, truth photons) of a case; ``cpt`` = OP_CHANNEL_PER_TRIG."""
    rng = np.random.default_rng(SEEDS[name])
    if name == "module0":              # threshold mode, all 96 channels, truth M = 2, several pulses, front padding
        ndet, nticks, M = 96, 9000, 2
        op = np.arange(ndet)
        pulses = [(300, 3000.0, np.arange(0, 48)), (1500, 2500.0, np.arange(48, 96)), (4200, 900.0, np.arange(0, 96)),
                  (4300, 3500.0, np.arange(12, 30)), (8100, 4000.0, np.arange(0, 96))]
        f8 = False
    elif name == "module0_sparse":     # a subset of the channels simulated (missing ones are filled in), no truth, float64 input
        M = 0
        op = np.sort(np.random.default_rng(3).choice(96 // cpt, 9, replace=False))[:, None] * cpt + np.arange(cpt)[None, :]
        op = op.ravel()
        ndet, nticks = len(op), 5003
        pulses = [(1900, 5000.0, np.arange(0, ndet)), (2400, 900.0, np.arange(0, 12)), (4950, 2000.0, np.arange(0, 12))]
        f8 = True
    else:                              # 2x2: beam trigger mode (one trigger at tick 0 for the first sub-batch)
        ndet, nticks, M = n_op_channel // 4, 3000, 1
        op = np.arange(ndet)
        pulses = [(200, 1500.0, np.arange(0, ndet)), (1700, 800.0, np.arange(6, 40))]
        f8 = False
    sig, tid, tph = waveforms(rng, ndet, nticks, pulses, M, f8)
    return sig, op.astype(np.int64), tid, tph


def load(name):
    return np.load(os.path.join(ROOT, "tests", "golden", "light_trigger_%s.npz" % name))


def consts_from_npz(z):
    C = {k[2:]: z[k] for k in z.files if k.startswith("c_")}
    for k in ("OP_CHANNEL_PER_TRIG", "LIGHT_TRIG_MODE", "LIGHT_NBIT", "N_OP_CHANNEL"):
        C[k] = int(C[k])
    for k in ("LIGHT_DIGIT_SAMPLE_SPACING", "LIGHT_TICK_SIZE", "MC_TRUTH_THRESHOLD"):
        C[k] = float(C[k])
    C["LIGHT_TRIG_WINDOW"] = tuple(float(x) for x in C["LIGHT_TRIG_WINDOW"])
    C["TPC_TO_MODULE"] = {int(a): int(b) for a, b in C["TPC_TO_MODULE"]}
    C["MODULE_TO_TPCS"] = {int(r[0]): [int(x) for x in r[1:]] for r in C["MODULE_TO_TPCS"]}
    return C


def thresholds(C, op):
    cpt = C["OP_CHANNEL_PER_TRIG"]
    thr_all = np.repeat(np.asarray(C["LIGHT_TRIG_THRESHOLD"])[..., np.newaxis], cpt, axis=-1).ravel()
    return thr_all[op].copy().reshape(-1, cpt)[..., 0]


# ---------------------------------------------------------------------------------------------------------
# light window extent (get_nticks / get_active_op_channel)
# ---------------------------------------------------------------------------------------------------------
LINC_DTYPE = np.dtype([("segment_id", "u4"), ("n_photons_det", "f4"), ("t0_det", "f4")])
EXTENT_CASES = ("lit", "dark", "one_entry", "negative_times", "empty")


def extent_inputs(case, ndet=96):
    rng = np.random.default_rng(sum(map(ord, case)))
    S = {"lit": 150, "dark": 40, "one_entry": 60, "negative_times": 80, "empty": 0}[case]
    li = np.zeros((S, ndet), dtype=LINC_DTYPE)
    li["segment_id"] = np.arange(S, dtype=np.uint32)[:, None]
    li["t0_det"] = (rng.random((S, ndet)) * 300 - (150 if case == "negative_times" else 0)).astype(np.float32)
    if case in ("lit", "negative_times"):
        nph = rng.random((S, ndet)) * 50
        nph[rng.random((S, ndet)) < 0.6] = 0
        nph[:, rng.random(ndet) < 0.4] = 0              # channels nobody lights
        li["n_photons_det"] = nph.astype(np.float32)
    elif case == "one_entry":
        li["n_photons_det"][S // 2, 17] = 3.5
    return li


def truth_inputs(case, ndet=96):
    """(true_track_id i8[ntrig, ndet, nsamples, M], true_photons f8[...]) with -1 holes; 'multi' has 3 triggers"""
    rng = np.random.default_rng(1000 + sum(map(ord, case)))
    nt, ns, M = {"single": (1, 40, 3), "multi": (3, 25, 2), "none": (2, 10, 2), "zero_triggers": (0, 10, 2)}[case]
    ids = rng.integers(0, 5000, (nt, ndet, ns, M)).astype(np.int64)
    ids[rng.random(ids.shape) < (1.0 if case == "none" else 0.85)] = -1
    ph = rng.random(ids.shape) * 20 - 2
    return ids, ph


TRUTH_CASES = ("single", "multi", "none", "zero_triggers")
