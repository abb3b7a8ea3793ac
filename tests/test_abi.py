"""The C-ABI library loads without a GPU and exports every symbol include/larndsim_b200.h declares; the
ctypes mirrors of the POD structs have the sizes the C compiler gives them; the product has no CPU fallback."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from larndsim_b200 import _abi, consts as lc, synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "larndsim_b200.h")


def _declared():
    txt = open(HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lsb_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    assert os.path.exists(_abi.LIB_PATH), "run python __graft_entry__.py first"
    lib = C.CDLL(_abi.LIB_PATH)
    names = _declared()
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing
    assert lib.lsb_abi_version() == 2


def test_no_extra_exports():
    out = subprocess.run(["nm", "-D", "--defined-only", _abi.LIB_PATH], capture_output=True, text=True).stdout
    exported = sorted(ln.split()[2] for ln in out.splitlines() if " T " in ln)
    assert exported == _declared()


def test_struct_sizes_match_c(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "larndsim_b200.h"\nint main(){printf("%zu %zu %zu %zu %zu\\n", sizeof(lsb_consts),'
                   ' sizeof(lsb_track_layout), sizeof(lsb_lut_layout), sizeof(lsb_linc_layout), sizeof(lsb_chain_result));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)])
    sizes = [int(x) for x in subprocess.check_output([str(exe)]).split()]
    assert sizes == [C.sizeof(_abi.Consts), C.sizeof(_abi.TrackLayout), C.sizeof(_abi.LutLayout), C.sizeof(_abi.LincLayout),
                     C.sizeof(_abi.ChainResult)]


def test_track_layout_by_name():
    L = _abi.track_layout(synth.segment_dtype)
    assert L.itemsize == 152
    assert L.offset[_abi.FIELDS.index("n_electrons")] == 56 and L.dtype[_abi.FIELDS.index("n_electrons")] == _abi.LSB_U32
    assert L.offset[_abi.FIELDS.index("t0")] == 96 and L.dtype[_abi.FIELDS.index("t0")] == _abi.LSB_F64
    L8 = _abi.track_layout(synth.test_dtype_f8)
    assert L8.dtype[_abi.FIELDS.index("pixel_plane")] == _abi.LSB_F64
    with pytest.raises(TypeError):
        _abi.track_layout(np.dtype("f4"))


def test_snapshots_and_consts():
    for name in lc.available_configs():
        c = lc.snapshot(lc.load_snapshot(name))
        assert c.n_tpc in (2, 8, 70) and c.n_time_ticks in (2001, 3201)
    c = lc.snapshot(lc.load_snapshot("module0"))
    assert (c.n_pixels[0], c.n_pixels[1]) == (140, 280) and abs(c.pixel_pitch - 0.4434) < 1e-9
    assert c.max_tracks_per_pixel == 50 and c.max_adc_values == 30


def test_kernel_launch_protocol_surface():
    from larndsim_b200 import quenching, drifting, pixels_from_track, detsim, fee, lightLUT, light_sim
    for mod, names in ((quenching, ["quench"]), (drifting, ["drift"]), (pixels_from_track, ["max_pixels", "get_pixels"]),
                       (detsim, ["time_intervals", "tracks_current_mc", "tracks_current", "get_track_pixel_map",
                                 "get_track_pixel_map2", "sum_pixel_signals"]), (fee, ["get_adc_values"]),
                       (lightLUT, ["calculate_light_incidence"]),
                       (light_sim, ["sum_light_signals", "calc_scintillation_effect", "calc_stat_fluctuations",
                                    "calc_light_detector_response"])):
        for n in names:
            k = getattr(mod, n)
            assert callable(k[1, 128]) and callable(k[(1, 1, 2), (1, 1, 64)]) and callable(k[1, 2, 0, 0])
    lc.load_snapshot("module0")
    assert pixels_from_track.id2pixel(20066) == (46, 143, 0)
    assert pixels_from_track.pixel2id(46, 143, 0) == 20066
    assert pixels_from_track.id2pixel(-1) == (139, 279, -1)             # Python floor semantics on padding ids


def test_product_has_no_cpu_fallback():
    """Without CUDA the kernels must raise, not compute on the host; nothing under the package imports oracle/."""
    import torch
    from larndsim_b200 import quenching
    lc.load_snapshot("module0")
    if not torch.cuda.is_available():
        seg = synth.cosmic_segments(4, lc.detector)
        with pytest.raises(RuntimeError):
            quenching.quench[1, 256](seg, 2)
        # the widened stages as well: selection / batching, light window extent, table lookup
        from larndsim_b200 import active_volume, light_sim
        from larndsim_b200.util import batching
        with pytest.raises(RuntimeError):
            active_volume.select_active_volume(seg, lc.detector.TPC_BORDERS)
        with pytest.raises(RuntimeError):
            next(batching.TPCBatcher(seg, seg, "event_id", tpc_batch_size=1, tpc_borders=lc.detector.TPC_BORDERS))
        import numpy as np
        inc = np.zeros((3, 4), dtype=[("n_photons_det", "f4"), ("t0_det", "f4")])
        with pytest.raises(RuntimeError):
            light_sim.get_nticks(inc)
        from larndsim_b200 import spill, chain
        with pytest.raises(RuntimeError):
            spill.SpillRunner(seg.dtype, synth.response_lut(lc.detector))
        with pytest.raises(RuntimeError):
            chain.Chain(seg.dtype, synth.response_lut(lc.detector))
    pkg = os.path.join(ROOT, "larndsim_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "larnd_oracle" not in txt and "oracle/" not in txt.replace("tests / oracle", ""), f
