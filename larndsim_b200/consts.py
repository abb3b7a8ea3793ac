"""Constant provider for the B200 kernels.

The reference freezes ``larndsim.consts.{detector,physics,light,sim}`` module globals into its
JIT-compiled kernels and ``importlib.reload``s the kernel modules whenever they change
(cli/simulate_pixels.py:459-464, 688-693).  Here constants are read AT CALL TIME from a provider
and passed to the CUDA kernels as one POD struct (``lsb_consts``), so no reload is needed -- but
the kernel modules survive ``importlib.reload`` all the same.

Provider resolution (first match wins):
  1. a package set with :func:`use` (e.g. ``use(larndsim.consts)``);
  2. ``larndsim.consts`` if the reference package has been imported by the caller -- this is the
     drop-in scenario: simulate_pixels.py keeps calling ``consts.load_properties`` itself;
  3. the namespaces of this module (``detector``, ``physics``, ``light``, ``sim``, ``units``),
     filled by :func:`load_snapshot` from ``configs/*.json`` (derived from the reference YAMLs by
     tools/make_config_snapshots.py; the reference's consts loader itself is reused unmodified,
     SURVEY.md section 2 row 10, and is not rebuilt here).
"""
import json
import os
import sys
import types

import numpy as np

from . import _abi

detector = types.SimpleNamespace()
physics = types.SimpleNamespace(BOX_ALPHA=0.93, BOX_BETA=0.207, BIRKS_Ab=0.800, BIRKS_kb=0.0486,
                                W_ION=23.6e-6, BOX=1, BIRKS=2)
light = types.SimpleNamespace()
sim = types.SimpleNamespace()
units = types.SimpleNamespace(e=1.0, mV=1e-9, ns=1.0, mus=1000.0, cm=10.0, mm=1.0)

#: pixels_from_track.MAX_NEIGHBOR_BACKTRACK_DISTANCE (pixels_from_track.py:11)
MAX_NEIGHBOR_BACKTRACK_DISTANCE = 4

_forced = None
_loaded = None
CONFIG_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "configs")


def use(pkg):
    """Force the provider (an object with .detector/.physics/.light/.sim/.units), or None."""
    global _forced
    _forced = pkg


def available_configs():
    return sorted(f[:-5] for f in os.listdir(CONFIG_DIR) if f.endswith(".json"))


def load_snapshot(name):
    """Fill this module's namespaces from ``configs/<name>.json`` (or a path)."""
    global _loaded
    path = name if os.path.exists(name) else os.path.join(CONFIG_DIR, name + ".json")
    with open(path) as f:
        snap = json.load(f)
    for ns, key in ((detector, "detector"), (light, "light"), (sim, "sim"), (physics, "physics"), (units, "units")):
        ns.__dict__.clear()
        for k, v in snap[key].items():
            if isinstance(v, list) and k != "PIXEL_CONNECTION_DICT":
                v = np.array(v)
            setattr(ns, k, v)
    detector.TIME_TICKS = np.linspace(detector.TIME_INTERVAL[0], detector.TIME_INTERVAL[1], int(detector.N_TIME_TICKS))
    detector.N_PIXELS = tuple(int(v) for v in detector.N_PIXELS)
    if hasattr(detector, "PIXEL_CONNECTION_DICT"):          # stored as rows [x, y, chip, channel]
        detector.PIXEL_CONNECTION_DICT = {(r[0], r[1]): (r[2], r[3]) for r in detector.PIXEL_CONNECTION_DICT}
    _loaded = name
    return sys.modules[__name__]


def provider():
    if _forced is not None:
        return _forced
    ref = sys.modules.get("larndsim.consts")
    if ref is not None and hasattr(ref, "detector"):
        return ref
    if _loaded is None:
        raise RuntimeError("no constants loaded: call larndsim_b200.consts.load_snapshot(<config>) "
                           "or import larndsim.consts and load_properties() first")
    return sys.modules[__name__]


def snapshot(p=None):
    """Read every constant the kernels use from the provider, now, into an ``lsb_consts``."""
    p = p or provider()
    d, ph, li, si, un = p.detector, p.physics, p.light, p.sim, p.units
    c = _abi.Consts()
    c.box_alpha, c.box_beta, c.birks_ab, c.birks_kb, c.w_ion = ph.BOX_ALPHA, ph.BOX_BETA, ph.BIRKS_Ab, ph.BIRKS_kb, ph.W_ION
    c.mode_box, c.mode_birks = int(ph.BOX), int(ph.BIRKS)
    c.e_field, c.lar_density, c.v_drift = d.E_FIELD, d.LAR_DENSITY, d.V_DRIFT
    c.electron_lifetime, c.long_diff, c.tran_diff = d.ELECTRON_LIFETIME, d.LONG_DIFF, d.TRAN_DIFF
    c.time_sampling, c.time_padding, c.time_window = d.TIME_SAMPLING, d.TIME_PADDING, d.TIME_WINDOW
    c.time_interval[0], c.time_interval[1] = float(d.TIME_INTERVAL[0]), float(d.TIME_INTERVAL[1])
    c.response_sampling, c.response_bin_size, c.pixel_pitch = d.RESPONSE_SAMPLING, d.RESPONSE_BIN_SIZE, d.PIXEL_PITCH
    c.n_time_ticks = int(len(d.TIME_TICKS))
    c.n_pixels[0], c.n_pixels[1] = int(d.N_PIXELS[0]), int(d.N_PIXELS[1])
    borders = np.ascontiguousarray(np.asarray(d.TPC_BORDERS, dtype=np.float64))
    ntpc = borders.shape[0]
    if ntpc > _abi.LSB_MAX_TPC:
        raise ValueError("too many TPCs (%d > %d)" % (ntpc, _abi.LSB_MAX_TPC))
    c.n_tpc = ntpc
    flat = borders.reshape(-1)
    for i in range(flat.size):
        c.tpc_borders[i] = flat[i]
    c.default_plane_index = int(d.DEFAULT_PLANE_INDEX)
    c.sampled_points = int(d.SAMPLED_POINTS)
    c.max_neighbor_backtrack_distance = int(MAX_NEIGHBOR_BACKTRACK_DISTANCE)
    c.discrimination_threshold = d.DISCRIMINATION_THRESHOLD
    c.adc_hold_delay, c.adc_busy_delay, c.reset_cycles, c.clock_cycle = d.ADC_HOLD_DELAY, d.ADC_BUSY_DELAY, d.RESET_CYCLES, d.CLOCK_CYCLE
    c.gain, c.buffer_risetime, c.v_cm, c.v_ref, c.v_pedestal = d.GAIN, d.BUFFER_RISETIME, d.V_CM, d.V_REF, d.V_PEDESTAL
    c.adc_counts = d.ADC_COUNTS
    c.reset_noise_charge, c.uncorrelated_noise_charge, c.discriminator_noise = d.RESET_NOISE_CHARGE, d.UNCORRELATED_NOISE_CHARGE, d.DISCRIMINATOR_NOISE
    c.unit_e, c.unit_mV, c.unit_ns, c.unit_mus = un.e, un.mV, un.ns, un.mus
    c.w_ph = getattr(li, "W_PH", 19.5e-6)
    c.scint_prescale = getattr(li, "SCINT_PRESCALE", 1)
    c.light_tick_size = getattr(li, "LIGHT_TICK_SIZE", 0.001)
    lw = getattr(li, "LIGHT_WINDOW", (1, 10))
    c.light_window[0], c.light_window[1] = float(lw[0]), float(lw[1])
    c.singlet_fraction = getattr(li, "SINGLET_FRACTION", 0.3)
    c.tau_s, c.tau_t = getattr(li, "TAU_S", 0.001), getattr(li, "TAU_T", 1.530)
    c.light_response_time = getattr(li, "LIGHT_RESPONSE_TIME", 0.055)
    c.light_oscillation_period = getattr(li, "LIGHT_OSCILLATION_PERIOD", 0.095)
    c.impulse_tick_size = getattr(li, "IMPULSE_TICK_SIZE", 0.001)
    c.sipm_response_model = int(getattr(li, "SIPM_RESPONSE_MODEL", 0))
    c.light_trig_mode = int(getattr(li, "LIGHT_TRIG_MODE", 0))
    c.enable_lut_smearing = int(bool(getattr(li, "ENABLE_LUT_SMEARING", False)))
    c.n_op_channel = int(getattr(li, "N_OP_CHANNEL", 0))
    c.min_step_size, c.mc_truth_threshold = si.MIN_STEP_SIZE, si.MC_TRUTH_THRESHOLD
    c.max_tracks_per_pixel, c.mc_sample_multiplier = int(si.MAX_TRACKS_PER_PIXEL), int(si.MC_SAMPLE_MULTIPLIER)
    c.max_adc_values = int(si.MAX_ADC_VALUES)
    return c
