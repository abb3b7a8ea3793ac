"""Derive compact constant snapshots from the reference's YAML configuration.

Runs ONLY in the build container: loads the detector / pixel-layout / simulation YAML files
through the reference's own ``larndsim.consts`` loaders (unmodified, from /root/reference) and
writes the *derived numbers* the kernels need (TPC borders, pixel grid, timing, FEE, light and
simulation constants) as JSON under ``larndsim_b200/configs/``.  The GPU box has no reference
tree; ``larndsim_b200.consts.load_snapshot(name)`` reads these files there.  In a real larnd-sim
installation the host layer reads ``larndsim.consts`` directly instead (consts.py).
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(__file__))
import refharness as rh  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "larndsim_b200", "configs")

DET = ["LAR_DENSITY", "E_FIELD", "V_DRIFT", "ELECTRON_LIFETIME", "LONG_DIFF", "TRAN_DIFF", "TEMPERATURE",
       "DRIFT_LENGTH", "TPC_BORDERS", "TIME_SAMPLING", "TIME_INTERVAL", "TIME_PADDING", "TIME_WINDOW",
       "SAMPLED_POINTS", "RESPONSE_SAMPLING", "RESPONSE_BIN_SIZE", "DEFAULT_PLANE_INDEX", "N_PIXELS",
       "N_PIXELS_PER_TILE", "PIXEL_PITCH", "DISCRIMINATION_THRESHOLD", "ADC_HOLD_DELAY", "ADC_BUSY_DELAY",
       "RESET_CYCLES", "CLOCK_CYCLE", "GAIN", "BUFFER_RISETIME", "V_CM", "V_REF", "V_PEDESTAL", "ADC_COUNTS",
       "RESET_NOISE_CHARGE", "UNCORRELATED_NOISE_CHARGE", "DISCRIMINATOR_NOISE", "MODULE_TO_TPCS",
       # readout tables of the packet builder (fee.export_to_hdf5)
       "CLOCK_RESET_PERIOD", "MODULE_TO_IO_GROUPS", "TILE_MAP", "TILE_ORIENTATIONS", "TILE_CHIP_TO_IO",
       "EVENT_RATE", "NON_BEAM_EVENT_GAP"]
LIGHT = ["LIGHT_SIMULATED", "ENABLE_LUT_SMEARING", "N_OP_CHANNEL", "OP_CHANNEL_EFFICIENCY", "OP_CHANNEL_TO_TPC",
         "SCINT_PRESCALE", "W_PH", "LIGHT_TICK_SIZE", "LIGHT_WINDOW", "SINGLET_FRACTION", "TAU_S", "TAU_T",
         "LIGHT_GAIN", "SIPM_RESPONSE_MODEL", "LIGHT_RESPONSE_TIME", "LIGHT_OSCILLATION_PERIOD",
         "IMPULSE_MODEL", "IMPULSE_TICK_SIZE", "LIGHT_TRIG_MODE", "OP_CHANNEL_PER_TRIG",
         "LIGHT_DIGIT_SAMPLE_SPACING", "LIGHT_NBIT", "LIGHT_TRIG_WINDOW", "TPC_TO_OP_CHANNEL", "LIGHT_TRIG_THRESHOLD",
         "LIGHT_DET_NOISE_SAMPLE_SPACING"]
SIM = ["BATCH_SIZE", "EVENT_BATCH_SIZE", "EVENT_SEPARATOR", "MAX_TRACKS_PER_PIXEL", "MIN_STEP_SIZE",
       "MC_SAMPLE_MULTIPLIER", "ASSOCIATION_COUNT_TO_STORE", "MAX_ADC_VALUES", "MAX_MC_TRUTH_IDS",
       "MC_TRUTH_THRESHOLD", "WRITE_BATCH_SIZE", "IS_SPILL_SIM", "SPILL_PERIOD", "MAX_EVENTS_PER_FILE"]
PHYS = ["BOX_ALPHA", "BOX_BETA", "BIRKS_Ab", "BIRKS_kb", "W_ION", "BOX", "BIRKS", "E_CHARGE"]
UNITS = ["e", "mV", "ns", "mus", "cm", "mm", "s"]

CONFIGS = {
    # name: (detector yaml, pixel layout, sim yaml, i_module)
    "module0": ("module0.yaml", "multi_tile_layout-2.3.16.yaml", "singles_sim.yaml", -1),
    "2x2": ("2x2_no_modvar.yaml", "multi_tile_layout-2.4.16.yaml", "2x2_NuMI_sim_no_modvar.yaml", -1),
    "2x2_mod2mod_variation_mod3": ("2x2.yaml", "multi_tile_layout-2.5.16.yaml", "2x2_NuMI_sim.yaml", 3),
    "2x2_mod2mod_variation_mod1": ("2x2.yaml", "multi_tile_layout-2.4.16.yaml", "2x2_NuMI_sim.yaml", 1),
    "ndlar": ("ndlar-module.yaml", "multi_tile_layout-3.0.40.yaml", "NDLAr_LBNF_sim.yaml", -1),
}


def jsonable(v):
    if isinstance(v, np.ndarray):
        return v.tolist()
    if isinstance(v, (np.floating, np.integer, np.bool_)):
        return v.item()
    if isinstance(v, (tuple, list)):
        return [jsonable(x) for x in v]
    if isinstance(v, dict):
        return {str(k): jsonable(x) for k, x in v.items()}
    return v


def main():
    import importlib
    rh.load_reference()
    from larndsim import consts
    os.makedirs(OUT, exist_ok=True)
    for name, (det, pix, sim, imod) in CONFIGS.items():
        for m in (consts.detector, consts.light, consts.sim):
            importlib.reload(m)
        rh.load_properties(det, pix, sim, imod)
        snap = {"source": {"detector_properties": det, "pixel_layout": pix, "simulation_properties": sim,
                           "i_module": imod, "generator": "tools/make_config_snapshots.py"},
                "detector": {k: jsonable(getattr(consts.detector, k)) for k in DET},
                "light": {k: jsonable(getattr(consts.light, k)) for k in LIGHT},
                "sim": {k: jsonable(getattr(consts.sim, k)) for k in SIM},
                "physics": {k: jsonable(getattr(consts.physics, k)) for k in PHYS},
                "units": {k: jsonable(getattr(consts.units, k)) for k in UNITS}}
        snap["detector"]["N_TIME_TICKS"] = int(len(consts.detector.TIME_TICKS))
        # {(x, y): (chip, channel)} -> rows [x, y, chip, channel] (JSON has no tuple keys)
        snap["detector"]["PIXEL_CONNECTION_DICT"] = sorted([int(k[0]), int(k[1]), int(v[0]), int(v[1])]
                                                           for k, v in consts.detector.PIXEL_CONNECTION_DICT.items())
        with open(os.path.join(OUT, name + ".json"), "w") as f:
            json.dump(snap, f, indent=0, separators=(",", ":"))
        print(name, "TPCs", consts.detector.TPC_BORDERS.shape[0], "N_PIXELS", consts.detector.N_PIXELS,
              "Tt", len(consts.detector.TIME_TICKS), "light", consts.light.LIGHT_SIMULATED, consts.light.N_OP_CHANNEL)


if __name__ == "__main__":
    main()
