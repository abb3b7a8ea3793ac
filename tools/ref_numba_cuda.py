#!/usr/bin/env python
"""Reference arm on the GPU: the UNMODIFIED reference kernels (``baseline/_ref/larndsim``, installed there by
``pip install --no-deps --target baseline/_ref /root/reference``; git-ignored, travels to the GPU box) compiled by
Numba-CUDA for the same B200 and timed kernel by kernel on the benchmark batch.

    python tools/ref_numba_cuda.py [--config module0] [--segments 10000] [--kind cosmic] [--out file.json]

The call sequence is cli/simulate_pixels.py:902-1102 with the reference's own launch geometries; CuPy is not in the image,
so the glue between the kernels (array creation, unique, pixel_index_map) is torch / NumPy and is NOT timed -- the number
reported is the sum of the reference's kernel times (CUDA events), which favours the reference.  Prints one JSON object;
``{"unavailable": "..."}`` with the exact error if Numba cannot drive this GPU.  Nothing of the product is on this path except
the synthetic input generator (larndsim_b200.synth, NumPy only).
"""
import argparse
import importlib
import json
import os
import sys
import time
import traceback
import types
from math import ceil

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")


def install_shims():
    """cupy / h5py / larpix are imported at module level by the reference but not used inside its kernels."""
    if "cupy" not in sys.modules:
        try:
            importlib.import_module("cupy")
        except ImportError:
            cp = types.ModuleType("cupy")
            cp.__dict__.update({k: getattr(np, k) for k in dir(np) if not k.startswith("__")})
            cp.get_array_module = lambda *a: np
            cp.asnumpy = lambda a: np.asarray(a)
            cuda = types.ModuleType("cupy.cuda")
            nvtx = types.ModuleType("cupy.cuda.nvtx")
            nvtx.RangePush = lambda *a, **k: None
            nvtx.RangePop = lambda *a, **k: None
            cuda.nvtx = nvtx
            cp.cuda = cuda
            sys.modules.update({"cupy": cp, "cupy.cuda": cuda, "cupy.cuda.nvtx": nvtx})
    if "h5py" not in sys.modules:
        try:
            importlib.import_module("h5py")
        except ImportError:
            sys.modules["h5py"] = types.ModuleType("h5py")
    if "larpix" not in sys.modules:
        try:
            importlib.import_module("larpix")
        except ImportError:
            lp = types.ModuleType("larpix")
            for sub, names in (("packet", ["Packet_v2", "TimestampPacket", "TriggerPacket", "SyncPacket", "PacketCollection"]),
                               ("key", ["Key"]), ("format", ["hdf5format"])):
                m = types.ModuleType("larpix." + sub)
                for n in names:
                    setattr(m, n, type(n, (), {}))
                setattr(lp, sub, m)
                sys.modules["larpix." + sub] = m
            sys.modules["larpix"] = lp


CONFIGS = {"module0": ("module0.yaml", "multi_tile_layout-2.3.16.yaml"),
           "2x2": ("2x2.yaml", "multi_tile_layout-2.4.16.yaml"),
           "ndlar": ("ndlar-module.yaml", "multi_tile_layout-3.0.40.yaml")}


def run(config, n_segments, kind, seed, rng_seed):
    if not os.path.isdir(os.path.join(REF, "larndsim")):
        return {"unavailable": "baseline/_ref/larndsim is missing (pip install --no-deps --target baseline/_ref /root/reference)"}
    install_shims()
    sys.path.insert(0, REF)
    sys.path.insert(0, ROOT)
    from numba import cuda
    if not cuda.is_available():
        return {"unavailable": "numba.cuda.is_available() is False"}
    import numba
    import torch
    from numba.cuda.random import create_xoroshiro128p_states
    from larndsim import consts
    det_yaml, pix_yaml = CONFIGS[config]
    p = os.path.join(REF, "larndsim")
    consts.load_properties(os.path.join(p, "detector_properties", det_yaml), os.path.join(p, "pixel_layouts", pix_yaml),
                           os.path.join(p, "simulation_properties", "singles_sim.yaml"))
    from larndsim import quenching, drifting, pixels_from_track, detsim, fee
    from larndsim.consts import detector, physics, sim
    from larndsim_b200 import synth
    if kind == "cosmic":
        tracks = synth.cosmic_segments(n_segments, detector, seed=seed)
    else:
        tracks = synth.beam_spill_segments(n_segments, detector, seed=seed)
    response = synth.response_lut(detector)
    dev = torch.device("cuda", 0)
    times = {}

    def timed(name, fn):
        e0, e1 = cuda.event(), cuda.event()
        cuda.synchronize()
        e0.record()
        fn()
        e1.record()
        e1.synchronize()
        times[name] = times.get(name, 0.0) + cuda.event_elapsed_time(e0, e1)

    def chain(tr, record):
        """cli/simulate_pixels.py:727-741 (quench, drift) + :902-1102 (the batch body)."""
        S = tr.shape[0]
        t = timed if record else (lambda name, fn: fn())
        d_tr = cuda.to_device(tr)
        TPB = 256
        BPG = max(ceil(S / TPB), 1)
        t("quench", lambda: quenching.quench[BPG, TPB](d_tr, physics.BIRKS))
        t("drift", lambda: drifting.drift[BPG, TPB](d_tr))
        h_tr = d_tr.copy_to_host()
        max_radius = ceil(max(h_tr["tran_diff"]) * 5 / detector.PIXEL_PITCH)
        TPB = 128
        BPG = max(ceil(S / TPB), 1)
        max_pixels = np.array([0])
        d_maxpix = cuda.to_device(max_pixels)
        t("max_pixels", lambda: pixels_from_track.max_pixels[BPG, TPB](d_tr, d_maxpix))
        max_pixels = d_maxpix.copy_to_host()
        max_neigh = int((2 * max_radius + 1) * max_pixels[0] + (1 + 2 * max_radius) * max_radius * 2)
        active = torch.full((S, int(max_pixels[0])), -1, dtype=torch.int32, device=dev)
        neigh = torch.full((S, max_neigh), -1, dtype=torch.int32, device=dev)
        nrad = torch.full((S, max_neigh), -1, dtype=torch.int32, device=dev)
        npl = torch.zeros(S, dtype=torch.float64, device=dev)
        t("get_pixels", lambda: pixels_from_track.get_pixels[BPG, TPB](d_tr, active, neigh, nrad, npl, max_radius))
        uniq = torch.unique(neigh.reshape(-1))
        uniq = uniq[uniq != -1].contiguous()
        U = int(uniq.shape[0])
        max_length = torch.zeros(1, dtype=torch.int64, device=dev)
        starts = torch.empty(S, dtype=torch.float64, device=dev)
        t("time_intervals", lambda: detsim.time_intervals[BPG, TPB](starts, max_length, d_tr))
        T = int(max_length.item())
        signals = torch.zeros((S, max_neigh, T), dtype=torch.float32, device=dev)
        TPB3 = (1, 1, 64)
        BPG3 = (max(ceil(S / 1), 1), max(ceil(max_neigh / 1), 1), max(ceil(T / 64), 1))
        rng_states = create_xoroshiro128p_states(int(np.prod(TPB3[:2]) * np.prod(BPG3[:2])), seed=rng_seed)
        d_resp = cuda.to_device(response)
        t("tracks_current_mc", lambda: detsim.tracks_current_mc[BPG3, TPB3](signals, neigh, d_tr, d_resp, rng_states))
        # pixel_index_map (glue; the reference loops over segments with CuPy compares): searchsorted here, untimed
        pim = torch.searchsorted(uniq, neigh.clamp(min=0)).to(torch.int64)
        pim = torch.where(neigh >= 0, pim, torch.full_like(pim, -1)).contiguous()
        K = int(sim.MAX_TRACKS_PER_PIXEL)
        tpm = torch.full((U, K), -1, dtype=torch.int64, device=dev)
        # the reference launches ceil(U/32) blocks of 32 without an index guard (detsim.py:578-580): pad U to a multiple of 32
        Upad = 32 * ceil(U / 32)
        uniq_p = torch.cat([uniq, torch.full((Upad - U,), -2, dtype=uniq.dtype, device=dev)])
        tpm_p = torch.full((Upad, K), -1, dtype=torch.int64, device=dev)
        t("get_track_pixel_map2", lambda: detsim.get_track_pixel_map2[Upad // 32, 32](tpm_p, uniq_p, neigh, nrad, int(nrad.max().item()) + 1))
        tpm.copy_(tpm_p[:U])
        Tt = len(detector.TIME_TICKS)
        ps = torch.zeros((U, Tt), dtype=torch.float64, device=dev)
        pts = torch.zeros((U, Tt, K), dtype=torch.float64, device=dev)
        oflow = torch.zeros(U, dtype=torch.float64, device=dev)
        t("sum_pixel_signals", lambda: detsim.sum_pixel_signals[BPG3, TPB3](ps, signals, starts, pim, tpm, pts, oflow))
        time_ticks = torch.linspace(0, detector.TIME_INTERVAL[1], Tt + 1, dtype=torch.float64, device=dev)
        A = int(sim.MAX_ADC_VALUES)
        integral = torch.zeros((U, A), dtype=torch.float64, device=dev)
        adc_ticks = torch.zeros((U, A), dtype=torch.float64, device=dev)
        cf = torch.zeros((U, A, K), dtype=torch.float64, device=dev)
        TPB = 128
        BPG = ceil(U / TPB)
        if TPB * BPG > len(rng_states):
            rng_states = create_xoroshiro128p_states(int(TPB * BPG), seed=rng_seed)
        thr = torch.full((U,), detector.DISCRIMINATION_THRESHOLD * consts.units.e, dtype=torch.float64, device=dev)
        t("get_adc_values", lambda: fee.get_adc_values[BPG, TPB](ps, pts, time_ticks, integral, adc_ticks, 0, rng_states, cf, thr))
        # fee.digitize is CuPy elementwise code in the reference (fee.py:499-515); not a Numba kernel, not timed
        hits = int((integral != 0).sum().item())
        return dict(S=S, P=max_neigh, U=U, T=T, Tt=Tt, hits=hits)

    t_c0 = time.perf_counter()
    chain(tracks[:min(64, len(tracks))].copy(), record=False)          # JIT compilation of every kernel
    compile_s = time.perf_counter() - t_c0
    w0 = time.perf_counter()
    shape = chain(tracks.copy(), record=True)
    wall_s = time.perf_counter() - w0
    kernel_ms = sum(times.values())
    cc = cuda.get_current_device().compute_capability
    return {"impl": "reference numba-cuda", "numba": numba.__version__, "device": cuda.get_current_device().name.decode()
            if isinstance(cuda.get_current_device().name, bytes) else str(cuda.get_current_device().name),
            "compute_capability": list(cc), "config": config, "kind": kind, "segments": int(shape["S"]), "shape": shape,
            "kernel_ms": times, "kernel_ms_total": kernel_ms, "segments_per_s_kernels_only": shape["S"] / (kernel_ms * 1e-3),
            "wall_s_including_untimed_glue": wall_s, "jit_compile_s": compile_s,
            "note": "sum of the reference's own Numba-CUDA kernel times (CUDA events), one pass, dense pixels_tracks_signals as the "
                    "reference allocates it; glue (unique, pixel_index_map, allocations) excluded"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="module0")
    ap.add_argument("--segments", type=int, default=10000)
    ap.add_argument("--kind", default="cosmic")
    ap.add_argument("--seed", type=int, default=12345)
    ap.add_argument("--rng-seed", type=int, default=1)
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    try:
        res = run(a.config, a.segments, a.kind, a.seed, a.rng_seed)
    except BaseException as exc:                      # record the exact failure: BASELINE.md quotes it
        res = {"unavailable": "%s: %s" % (type(exc).__name__, str(exc)[:2000]), "traceback": traceback.format_exc()[-4000:]}
    txt = json.dumps(res)
    if a.out:
        with open(a.out, "w") as f:
            f.write(txt + "\n")
    print(txt, flush=True)


if __name__ == "__main__":
    main()
