"""The drop-in modules keep the reference's call surface: every kernel / function the batch loop calls
(cli/simulate_pixels.py:727-1205) has the same positional parameter list as the reference's own definition.  Runs only where
the reference tree is present (the build container: /root/reference, or the installed copy under baseline/_ref)."""
import importlib
import inspect
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = next((p for p in (os.environ.get("LARNDSIM_REFERENCE", "/root/reference"), os.path.join(ROOT, "baseline", "_ref"))
            if os.path.isdir(os.path.join(p, "larndsim"))), None)

#: (module, attribute) the reference's loop calls; kernels are Numba dispatchers (``.py_func``), the rest plain functions
SURFACE = [
    ("quenching", "quench"), ("drifting", "drift"),
    ("pixels_from_track", "max_pixels"), ("pixels_from_track", "get_pixels"), ("pixels_from_track", "pixel2id"), ("pixels_from_track", "id2pixel"),
    ("detsim", "time_intervals"), ("detsim", "tracks_current_mc"), ("detsim", "tracks_current"), ("detsim", "get_track_pixel_map"),
    ("detsim", "get_track_pixel_map2"), ("detsim", "sum_pixel_signals"), ("detsim", "get_pixel_coordinates"),
    ("fee", "get_adc_values"), ("fee", "digitize"), ("fee", "export_to_hdf5"), ("fee", "export_sync_to_hdf5"),
    ("fee", "export_timestamp_trigger_to_hdf5"), ("fee", "gen_event_times"), ("fee", "rotate_tile"), ("fee", "get_trig_io"),
    ("lightLUT", "calculate_light_incidence"),
    ("light_sim", "get_nticks"), ("light_sim", "get_active_op_channel"), ("light_sim", "sum_light_signals"),
    ("light_sim", "calc_scintillation_effect"), ("light_sim", "calc_stat_fluctuations"), ("light_sim", "calc_light_detector_response"),
    ("light_sim", "gen_light_detector_noise"), ("light_sim", "get_triggers"), ("light_sim", "digitize_signal"), ("light_sim", "sim_triggers"),
    ("light_sim", "zero_suppress_waveform_truth"),
    ("active_volume", "select_active_volume"),
]


def _params(obj):
    fn = getattr(obj, "py_func", None) or obj
    return [(p.name, p.kind) for p in inspect.signature(fn).parameters.values() if not p.name.startswith("_")]


@pytest.fixture(scope="module")
def reference():
    if REF is None:
        pytest.skip("reference tree not present")
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    os.environ["LARNDSIM_REFERENCE"] = REF
    import refharness as rh
    rh.REF_ROOT = REF
    return rh.load_reference()


@pytest.mark.parametrize("module,name", SURFACE)
def test_same_positional_parameters(reference, module, name):
    ref_mod = reference.get(module) or importlib.import_module("larndsim." + module)
    ours_mod = importlib.import_module("larndsim_b200." + module)
    ref_obj, our_obj = getattr(ref_mod, name), getattr(ours_mod, name)
    ref_p, our_p = _params(ref_obj), _params(our_obj)
    positional = (inspect.Parameter.POSITIONAL_ONLY, inspect.Parameter.POSITIONAL_OR_KEYWORD)
    ref_names = [n for n, k in ref_p if k in positional]
    our_names = [n for n, k in our_p if k in positional]
    # same count and order; names must agree too (keyword calls such as export_to_hdf5(..., i_mod=i_mod) exist in the CLI)
    assert our_names[:len(ref_names)] == ref_names, (module, name, ref_names, our_names)
    # anything a drop-in adds must be optional
    extra = [p for p in inspect.signature(getattr(our_obj, "py_func", None) or our_obj).parameters.values()
             if p.name not in ref_names and not p.name.startswith("_")]
    assert all(p.default is not inspect.Parameter.empty or p.kind in (inspect.Parameter.VAR_KEYWORD, inspect.Parameter.VAR_POSITIONAL) for p in extra), extra


def test_batcher_signature(reference):
    from larndsim.util import batching as ref_b
    from larndsim_b200.util import batching as our_b
    ref_n = list(inspect.signature(ref_b.TPCBatcher.__init__).parameters)
    our_n = list(inspect.signature(our_b.TPCBatcher.__init__).parameters)
    assert our_n[:len(ref_n)] == ref_n


def test_kernel_launch_protocol():
    """kernel[griddim, blockdim](*args) and kernel[grid, block, stream, shmem] are accepted (Numba dispatcher syntax)"""
    from larndsim_b200 import quenching
    k = quenching.quench
    assert callable(k[1, 128]) and callable(k[(1, 1), (1, 1, 64), 0, 0])
    with pytest.raises(ValueError):
        k[1, 2, 3, 4, 5]
