#!/usr/bin/env python
"""Benchmark of the charge-readout path (BASELINE.json metric: segments/s quench -> ADC at 1/2/4/8 B200).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (every N): ONE fixed synthetic ND-LAr beam spill -- ndlar config (70 TPCs), 1e6 segments in 4 events, seed 12345
(BASELINE.json configs[4]) -- run the way the reference runs a file (cli/simulate_pixels.py:667-1117 +
save_results -> fee.export_to_hdf5): active-volume cut -> quench -> drift -> (event, TPC pair) batches of
larndsim/util/batching.py -> per batch get_pixels -> tracks_current_mc -> sum_pixel_signals -> get_adc_values -> digitize ->
LArPix packets + mc_packets_assn rows.  A step = the whole spill.  With N ranks (torchrun, one per GPU) the 140 batches are
assigned to the ranks longest-first, there is no collective inside the chain, and every rank's packets go to rank 0 over
NCCL where they are put into file order: STRONG scaling, `value` = 1e6 segments / max-over-ranks device time per step.

`value`: records resident in HBM when the timed region starts, packets left in rank 0's HBM.  `e2e`: the same call with the
records in pinned host memory and the packets, truth rows and updated records copied back to host arrays (H2D/D2H inside).

`--impl reference` times the CPU implementation of the same path: the reference's kernels are Numba Python (CUDA simulator:
~1e3 s per segment), so this arm runs the C/OpenMP restatement pinned to them (oracle/larnd_oracle.c) with every host core on
a bounded sample of the same spill.  The reference's own Numba-CUDA build on the same B200 is timed by
tools/ref_numba_cuda.py (block `reference_numba_cuda` of our line).
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

# the driver runs the defaults; the other BASELINE.json configs are measured by setting these (e.g. configs[2]:
# LSB_BENCH_CONFIG=2x2 LSB_BENCH_SEGMENTS=200000 LSB_BENCH_TPC_BATCH=2 python bench.py --gpus 4)
CONFIG = os.environ.get("LSB_BENCH_CONFIG", "ndlar")
N_SEGMENTS = int(os.environ.get("LSB_BENCH_SEGMENTS", 1_000_000))
N_EVENTS = int(os.environ.get("LSB_BENCH_EVENTS", 4))
TPC_BATCH = int(os.environ["LSB_BENCH_TPC_BATCH"]) if os.environ.get("LSB_BENCH_TPC_BATCH") else None
SEED = 12345
RAND_SEED = 1
METRIC = "segments/s quench->ADC (%s config, synthetic beam spill of %.0e segments in %d events, (event, TPC pair) batches)" % (CONFIG, N_SEGMENTS, N_EVENTS)
WORKLOAD = (CONFIG + " config, synthetic full beam spill (%d segments, %d events, seed %d), charge readout quench->drift->get_pixels->"
            "tracks_current_mc->sum_pixel_signals->get_adc_values->digitize->packets, noise on, one fixed spill partitioned by "
            "(event, TPC pair) over the GPUs") % (N_SEGMENTS, N_EVENTS, SEED)


def make_spill():
    """(constants module, records, response table); the spill is cached under /tmp for the other arm / ranks"""
    from larndsim_b200 import consts as lc, synth
    mod = lc.load_snapshot(CONFIG)
    path = "/tmp/lsb_bench_spill_%s_%d_%d_%d.npy" % (CONFIG, N_SEGMENTS, N_EVENTS, SEED)
    tracks = None
    if os.path.exists(path):
        try:
            tracks = np.load(path)
            if tracks.dtype != synth.segment_dtype or len(tracks) != N_SEGMENTS:
                tracks = None
        except Exception:
            tracks = None
    if tracks is None:
        tracks = synth.beam_spill_segments(N_SEGMENTS, mod.detector, seed=SEED, n_events=N_EVENTS)
        tracks["segment_id"] = np.arange(len(tracks))
        tracks["file_traj_id"] = tracks["traj_id"]
        try:
            tmp = path + ".%d.tmp.npy" % os.getpid()
            np.save(tmp, tracks)
            os.replace(tmp, path)
        except OSError:
            pass
    return mod, tracks, synth.response_lut(mod.detector)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed regions (B200_PROFILING.md recipe).  nvidia-smi needs a moment to
    start, so the sampler runs from before the warm-up; `begin()` / `end()` bracket the timed regions (device-resident and e2e)
    and only samples taken inside them are reported (the nearest ones if a region is shorter than the sampling period)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx = gpu_index
        self.lines = []               # (arrival time, text)
        self.windows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thr = threading.Thread(target=self._read, daemon=True)
            self.thr.start()
        except OSError:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.lines.append((time.perf_counter(), ln.strip()))

    def begin(self):
        self._t0 = time.perf_counter()

    def end(self):
        self.windows.append((self._t0, time.perf_counter()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.1)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        parsed = []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for t, ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                parsed.append((t, float(f[1]), float(f[2]), [nm for nm, val in zip(names, f[5:9]) if val.lower().startswith("active")]))
            except ValueError:
                continue
        inside = [p for p in parsed if any(a - 0.03 <= p[0] <= b + 0.03 for a, b in self.windows)]
        where = "inside the timed regions"
        if not inside and parsed and self.windows:
            mid = 0.5 * (self.windows[0][0] + self.windows[-1][1])
            inside = sorted(parsed, key=lambda p: abs(p[0] - mid))[:3]
            where = "nearest to the timed regions"
        sm = [p[1] for p in inside]
        reasons = sorted({r for p in inside for r in p[3]})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(p[2] for p in inside) if inside else None,
                "reasons": reasons, "samples": len(sm), "sampled": where, "samples_whole_run": len(parsed)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


# ------------------------------------------------------------------------------------------------------------------------
# CPU arm: the pinned C/OpenMP restatement (oracle/) on a bounded sample of the same spill
# ------------------------------------------------------------------------------------------------------------------------
def use_all_host_threads():
    """torchrun exports OMP_NUM_THREADS=1: the CPU arm wants every core (set before libgomp initialises, and again through its API)"""
    n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)
    try:
        C.CDLL("libgomp.so.1").omp_set_num_threads(C.c_int(n))
    except OSError:
        pass
    return n


def cpu_sample(tracks, mod, n_sample):
    """the first `n_sample` segments of the first (event, TPC pair) batch of the spill that holds that many (else of the largest
    batch), as the loop would hand them to the chain"""
    det = mod.detector
    b = np.asarray(det.TPC_BORDERS, dtype=np.float64)
    lo, hi = np.minimum(b[:, :, 0], b[:, :, 1]), np.maximum(b[:, :, 0], b[:, :, 1])
    first = np.full(len(tracks), -1, dtype=np.int64)
    for tpc in range(b.shape[0] - 1, -1, -1):                  # lowest TPC index that contains the start or the end point
        for sfx in ("start", "end"):
            ins = np.ones(len(tracks), dtype=bool)
            for k, ax in enumerate("xyz"):
                v = tracks["%s_%s" % (ax, sfx)].astype(np.float64)
                ins &= (v > lo[tpc, k]) & (v < hi[tpc, k])
            first[ins] = tpc
    events = np.unique(tracks["event_id"])
    n_pairs = (b.shape[0] + 1) // 2
    unit = np.searchsorted(events, tracks["event_id"]) * n_pairs + first // 2
    unit[first < 0] = -1
    counts = np.bincount(unit[unit >= 0], minlength=len(events) * n_pairs)
    big = np.nonzero(counts >= n_sample)[0]
    u = int(big[0]) if len(big) else int(np.argmax(counts))
    return np.ascontiguousarray(tracks[unit == u][:n_sample])


def cpu_chain(sub, response, rng_seed=1):
    """The CPU restatement of the chain on one batch; returns (seconds, hits)."""
    import helpers as h
    sub = sub.copy()
    orc = h.Oracle()
    t0 = time.perf_counter()
    front = h.oracle_front(sub, orc, quench_mode=orc.c.mode_birks)
    S, P_ = front["neigh"].shape
    n_rng = max(S * P_, 128 * ((len(front["uniq"]) + 127) // 128))
    states = h.rng_states(n_rng, rng_seed)
    sig = orc.tracks_current_mc(sub, front["neigh"], front["T"], response, states, 0)
    back = h.oracle_back(orc, front, sig, states)
    dt = time.perf_counter() - t0
    return dt, int((back["digit"] > orc.digitize(np.zeros(1))[0]).sum())


def run_reference(args, rank, world):
    if rank != 0:
        return                                     # rank 0 alone runs and prints; the others exit 0 without work
    cores = use_all_host_threads()
    mod, tracks, response = make_spill()
    # bounded sample: sized from a calibration run so that the K timed steps take about 90 s on this host
    probe = cpu_sample(tracks, mod, 96)
    cpu_chain(probe[:16], response)
    t_probe, _ = cpu_chain(probe, response)
    n_auto = int(len(probe) / t_probe * 90.0 / max(args.steps, 1))
    n_sample = int(os.environ.get("LSB_BENCH_CPU_SAMPLE", max(128, min(4096, n_auto))))
    sub = cpu_sample(tracks, mod, n_sample)
    for _ in range(args.warmup):
        cpu_chain(sub[:32], response)
    times = []
    for _ in range(args.steps):
        dt, hits = cpu_chain(sub, response)
        times.append(dt)
    total = sum(times)
    value = len(sub) * args.steps / total
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "segments/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "sample": "per step the first %d segments of the spill's first (event, TPC pair) batch" % len(sub)},
            "cpu_baseline": {"value": value, "unit": "segments/s", "cores": cores, "kind": "port",
                             "sample": "first %d segments of the first batch of the spill per step; C/OpenMP restatement pinned to the "
                                       "reference's golden vectors (the reference itself is Numba Python: ~1e3 s per segment under "
                                       "the CUDA simulator)" % len(sub)},
            "e2e": {"value": value, "unit": "segments/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def parse_profile(txt):
    out = {}
    for ln in txt.strip().split("\n"):
        f = ln.split()
        if len(f) == 3:
            out[f[0]] = (int(f[1]), float(f[2]))
    return out


def profile_session(lib, fn):
    from larndsim_b200 import _launch as ll
    import torch
    torch.cuda.synchronize()
    lib.lsb_profile_begin(ll.stream())
    fn()
    torch.cuda.synchronize()
    buf = C.create_string_buffer(1 << 16)
    lib.lsb_profile_end(buf, C.c_int64(len(buf)))
    return parse_profile(buf.value.decode())


FP32_LANES = 148 * 128                                   # B200: 148 SMs x 128 FP32 lanes (no tensor cores on this path)


def kernel_rooflines(prof, steps, shape, peak_hbm, sm_mhz):
    """SURVEY.md 8(d) algorithmic work per kernel x the units of the profiled launches / measured launch time.
    `shape`: sums over the profiled batches -- n_fma, n_samples, pair_ticks (valid pairs x T), pixel_ticks (U x Tt), U, S, P."""
    fp32_peak = FP32_LANES * 2 * sm_mhz * 1e6 / 1e12
    A, K = shape["A"], shape["K"]
    alg = {
        # 2 flop per (sample, tick) pair that passes every test: the table FMA
        "k_mc_accumulate": ("fp32", 2.0 * shape["n_fma"]),
        # 4 B signals read per valid (pair, tick) + 8 B per pixel tick written (sparse form: the reference's 8 B slot write per
        # (pair, tick) into pixels_tracks_signals is not materialised, so it is not counted as work done)
        "k_sum_pixel_signals": ("hbm", 4.0 * shape["pair_ticks"] + 8.0 * shape["pixel_ticks"]),
        # get_adc_values: one pass over pixels_signals + the noise streams + the hit tables
        "k_fee_trigger": ("hbm", 8.0 * shape["pixel_ticks"] + 2 * 8.0 * shape["U"] * A),
        "k_fee_fir_pre": ("hbm", 2 * 8.0 * shape["pixel_ticks"]),
        # noise stream of every pixel: two float32 normals per tick written once (+ hold-delay draws)
        "k_fee_rng_chunks": ("hbm", 2 * 4.0 * shape["pixel_ticks"]),
        "k_fee_fractions_sparse": ("hbm", 4.0 * shape["pair_ticks"] + 8.0 * shape["U"] * A * K),
        "k_mc_sampler": ("fp32", 330.0 * shape["n_samples"]),
        "k_mc_uniforms": ("hbm", 24.0 * shape["n_samples"]),
        "k_mc_sort": ("hbm", 4.0 * shape["n_samples"] + 16.0 * shape["n_samples"] / 3.0),
    }
    out = {}
    for name, (bound, work) in alg.items():
        if name not in prof or prof[name][1] <= 0:
            continue
        cnt, ms = prof[name]
        if bound == "hbm":
            ach = work / (ms * 1e-3) / 1e9
            out[name] = {"bound": "hbm", "algorithmic_bytes": work, "ms": ms, "launches": cnt, "achieved": ach, "peak": peak_hbm,
                         "unit": "GB/s", "frac": ach / peak_hbm}
        else:
            ach = work / (ms * 1e-3) / 1e12
            out[name] = {"bound": "fp32", "algorithmic_flops": work, "ms": ms, "launches": cnt, "achieved": ach, "peak": fp32_peak,
                         "unit": "TFLOP/s", "frac": ach / fp32_peak}
    return out


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")       # keep NCCL's banner off stdout: rank 0 prints ONE JSON line
        import datetime
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=180))
    from larndsim_b200 import _launch as ll, chain as lchain, consts as lc, spill as lspill
    lib = ll.lib()
    lib.lsb_profile_end.restype = C.c_int64
    lib.lsb_launch_count.restype = C.c_int64
    mod, tracks, response = make_spill()
    S = len(tracks)
    itemsize = tracks.dtype.itemsize
    events = np.unique(tracks["event_id"])
    snap = lc.snapshot()
    A, K, Tt = int(snap.max_adc_values), int(snap.max_tracks_per_pixel), int(snap.n_time_ticks)
    depth = int(os.environ.get("LSB_BENCH_DEPTH", 4))
    runner = lspill.SpillRunner(tracks.dtype, response, depth=depth, tpc_batch_size=TPC_BATCH)
    raw = torch.from_numpy(tracks.view(np.uint8).reshape(-1).copy())
    pinned_in = raw.pin_memory()
    n_total = args.warmup + args.steps
    dev_copies = [ll.DeviceRecords(dtype=tracks.dtype, n=S, buf=pinned_in.cuda()) for _ in range(n_total)]

    def sync():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        t = torch.tensor([ms], device="cuda", dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    # ---------------- value: records resident in HBM, packets left in rank 0's HBM ----------------
    for i in range(args.warmup):
        runner.simulate(dev_copies[i], events=events, rand_seed=RAND_SEED, host_output=False)
    sync()
    launches0 = lib.lsb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.begin()
    e0.record()
    for i in range(args.steps):
        last = runner.simulate(dev_copies[args.warmup + i], events=events, rand_seed=RAND_SEED, host_output=False)
    e1.record()
    sync()
    sampler.end()
    launches = lib.lsb_launch_count() - launches0
    ms_own = e0.elapsed_time(e1)
    ms_max = max_over_ranks(ms_own)
    value = S * args.steps / (ms_max * 1e-3)
    st = last.stats
    per_rank = [{"ms_per_step": ms_own / args.steps, "units": st["n_units_here"], "segments": st["n_segments_here"],
                 "packets": st["n_packets_here"]}]
    if world > 1:
        v = torch.tensor([ms_own / args.steps, st["n_units_here"], st["n_segments_here"], st["n_packets_here"]], device="cuda", dtype=torch.float64)
        allv = [torch.zeros_like(v) for _ in range(world)]
        dist.all_gather(allv, v)
        per_rank = [{"ms_per_step": float(a[0]), "units": int(a[1]), "segments": int(a[2]), "packets": int(a[3])} for a in allv]
    del dev_copies
    # ---------------- e2e: pinned host records in, packets + truth rows + updated records in host arrays out ----------------
    host_np = pinned_in.numpy().view(tracks.dtype)
    for _ in range(2):
        out = runner.simulate(host_np, events=events, rand_seed=RAND_SEED, host_output=True, return_tracks=True)
    sync()
    sampler.begin()
    e0.record()
    for i in range(args.steps):
        out = runner.simulate(host_np, events=events, rand_seed=RAND_SEED, host_output=True, return_tracks=True)
    e1.record()
    sync()
    sampler.end()
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = max_over_ranks(e0.elapsed_time(e1))
    e2e_value = S * args.steps / (ms_e2e * 1e-3)
    if world > 1:                 # last collective of the job: rank 0 continues alone (per-kernel pass, secondary blocks)
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    n_packets = int(out.n_packets)
    n_data = int((out.packets["packet_type"] == 0).sum())
    h2d = S * itemsize
    d2h = n_packets * (32 + runner.assn_dtype.itemsize) + out.n_segments * itemsize
    unit_sizes = out.unit_sizes
    sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
    peak, peak_src = measured_peaks()

    # ---------------- per-kernel times: rank 0's share of the spill, one unit at a time on one stream ----------------
    roofline = None
    kernels = None
    by_kernel = None
    try:
        prof_runner = lspill.SpillRunner(tracks.dtype, response, depth=1, single_rank=True, tpc_batch_size=TPC_BATCH)      # the other ranks are done: no collective here
        prof_runner.set_serial(True)
        d_prof = ll.DeviceRecords(dtype=tracks.dtype, n=S, buf=pinned_in.cuda())
        prof_runner.simulate(ll.DeviceRecords(dtype=tracks.dtype, n=S, buf=pinned_in.cuda()), events=events, rand_seed=RAND_SEED, host_output=False)
        holder = {}

        def prof_pass():
            holder["out"] = prof_runner.simulate(d_prof, events=events, rand_seed=RAND_SEED, host_output=False)
        s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        s0.record()
        prof = profile_session(lib, prof_pass)
        s1.record()
        torch.cuda.synchronize()
        ms_serial = s0.elapsed_time(s1)
        pst = holder["out"].stats
        kern = [(k, v) for k, v in prof.items() if not k.startswith("(")]
        total_kernel_ms = sum(v[1] for _, v in kern)
        top_name, (top_cnt, top_ms) = max(kern, key=lambda kv: kv[1][1])
        shape = {"n_fma": pst["n_fma"], "n_samples": pst["n_samples"], "pair_ticks": pst["pair_ticks"], "pixel_ticks": pst["pixel_ticks"],
                 "U": pst["n_unique_pixels"], "A": A, "K": K}
        by_kernel = kernel_rooflines(prof, 1, shape, peak, sm_mhz)
        traffic, ncu = None, {}
        tpath = os.path.join(ROOT, "profiles", "r02_traffic.json")
        if os.path.exists(tpath):
            with open(tpath) as f:
                tj = json.load(f)
            ncu = tj.get(top_name, {})
            traffic = ncu.get("dram_bytes_per_launch")
        rk = by_kernel.get(top_name, {})
        roofline = {"kernel": top_name, "bound": rk.get("bound"), "achieved": rk.get("achieved"), "peak": rk.get("peak"), "unit": rk.get("unit"),
                    "frac": rk.get("frac"), "traffic": traffic,
                    "traffic_source": ((ncu.get("_source") or tj.get("_source")) if traffic else None) if os.path.exists(tpath) else None,
                    "timed": "CUDA events on the launching stream around every launch of the kernel: the whole spill on rank 0, one "
                             "batch at a time on one stream (%d launches)" % top_cnt,
                    "share_of_kernel_time": top_ms / total_kernel_ms, "ms_per_launch": top_ms / max(top_cnt, 1), "launches": top_cnt,
                    "peak_source": "148 SM x 128 FP32 lanes x 2 x %.0f MHz (non-tensor FP32: nothing on the path is a contraction)" % sm_mhz
                                   if rk.get("bound") == "fp32" else peak_src,
                    "work": "2 flop x N_fma; N_fma = (sample, tick) pairs that pass every test of detsim.py:299,333,341-344, counted on the "
                            "device by the sampler (%d in these launches)" % pst["n_fma"] if top_name == "k_mc_accumulate" else "SURVEY.md 8(d)"}
        if top_name == "k_mc_accumulate":
            t_s = top_ms * 1e-3
            onchip = 4.0 * pst["n_fma"] / t_s / 1e12
            roofline["onchip"] = {"what": "table words the algorithmic formulation reads from L1 / shared memory: 4 B x N_fma", "achieved": onchip,
                                  "peak": 148 * 128 * sm_mhz * 1e6 / 1e12, "unit": "TB/s", "frac": onchip / (148 * 128 * sm_mhz * 1e6 / 1e12),
                                  "ncu_l1tex_throughput_pct": ncu.get("l1tex_throughput_pct"), "ncu_issue_active_pct": ncu.get("issue_active_pct"),
                                  "ncu_lts_throughput_pct": ncu.get("lts_throughput_pct"), "ncu_l2_to_l1_TBs": ncu.get("l2_to_l1_TBs"),
                                  "ncu_note": ncu.get("_source")}
            l2p = os.path.join(ROOT, "profiles", "r02_l2_bandwidth.json")
            if ncu.get("l2_to_l1_bytes_per_fma") and os.path.exists(l2p):
                with open(l2p) as f:
                    l2peak = json.load(f)["l2_read_GBs"] / 1e3
                l2ach = ncu["l2_to_l1_bytes_per_fma"] * pst["n_fma"] / t_s / 1e12
                roofline["l2"] = {"what": "table bytes that travel L2 -> L1: (lts__t_sectors_srcunit_tex_op_read x 32 B / N_fma of the ncu capture, "
                                          "profiles/r02_traffic.json) x N_fma of this run / the kernel's time in this run",
                                  "achieved": l2ach, "peak": l2peak, "unit": "TB/s", "frac": l2ach / l2peak,
                                  "peak_source": "tools/l2_bandwidth.cu on this pool's B200 (LDG.128 stream over an L2-resident buffer, profiles/r02_l2_bandwidth.json)"}
            mc_ms = sum(prof[k][1] for k in ("k_mc_pairs", "k_mc_uniforms", "k_mc_sampler", "k_mc_sort", "k_mc_accumulate", "k_mc_fused") if k in prof)
            flops = 2.0 * pst["n_fma"] + 330.0 * pst["n_samples"]
            fp32_peak = FP32_LANES * 2 * sm_mhz * 1e6 / 1e12
            roofline["tracks_current_mc_stage"] = {"algorithmic_flops": flops, "ms": mc_ms, "achieved_tflops": flops / (mc_ms * 1e-3) / 1e12,
                                                   "peak_tflops": fp32_peak, "frac": flops / (mc_ms * 1e-3) / 1e12 / fp32_peak,
                                                   "work": "SURVEY 8(d): 2 N_fma + 330 N_sp (N_sp = %d sample points)" % pst["n_samples"]}
        kernels = {k: {"launches": v[0], "ms": v[1], "share": v[1] / total_kernel_ms} for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:14]}
        kernels["_total_kernel_ms"] = total_kernel_ms
        kernels["_wall_ms_of_the_profiled_pass"] = ms_serial
        prof_runner.close()
        del d_prof
    except Exception as exc:                                   # extras never lose the headline line
        roofline = {"error": repr(exc)}

    # ---------------- secondary workload: BASELINE configs[1], module0 1e4 cosmic segments per batch (round-1 headline) ----------------
    module0_block = None
    try:
        module0_block = bench_module0(lib, ll, lchain, lc, peak, sm_mhz, steps=max(3, min(args.steps, 10)))
    except Exception as exc:
        module0_block = {"error": repr(exc)}
    lc.load_snapshot(CONFIG)
    try:        # the same kernel on the round-1 workload, for comparison (coarser table: 4.1 instead of 1.75 samples per group record)
        m0 = module0_block["roofline_by_kernel"]["k_mc_accumulate"]
        t_s = m0["ms"] * 1e-3
        l1_peak = 148 * 128 * sm_mhz * 1e6 / 1e12
        roofline["same_kernel_on_module0_1e4"] = {"frac_of_fp32_peak": m0["frac"], "achieved_tflops": m0["achieved"],
                                                  "onchip_TBs": 2.0 * m0["algorithmic_flops"] / t_s / 1e12,
                                                  "onchip_frac": 2.0 * m0["algorithmic_flops"] / t_s / 1e12 / l1_peak}
    except Exception:
        pass
    # ---------------- light path (BASELINE configs[3]) ----------------
    light_block = None
    try:
        import bench_light
        light_block = bench_light.run(peak, steps=3)
    except Exception as exc:
        light_block = {"error": repr(exc)}
    lc.load_snapshot(CONFIG)
    # ---------------- CPU baseline (bounded sample) + the reference's own Numba-CUDA build on this GPU ----------------
    cpu = None
    ref_cuda = None
    if world == 1 and not os.environ.get("LSB_BENCH_NO_CPU"):
        cores = use_all_host_threads()
        sub = cpu_sample(tracks, mod, 2048)
        cpu_chain(sub[:16], response)
        cpu_s, _ = cpu_chain(sub, response)
        cpu = {"value": len(sub) / cpu_s, "unit": "segments/s", "cores": cores, "kind": "port",
               "sample": "first %d segments of the spill's first (event, TPC pair) batch, %.1f s; C/OpenMP restatement of the reference "
                         "kernels (oracle/), pinned to the reference's golden vectors" % (len(sub), cpu_s)}
        try:
            runner.close()
            torch.cuda.empty_cache()
            p = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ref_numba_cuda.py"), "--config", "module0", "--segments", "10000"],
                               capture_output=True, text=True, timeout=300)
            ref_cuda = json.loads(p.stdout.strip().split("\n")[-1])
            ref_cuda.pop("traceback", None)
        except Exception as exc:
            ref_cuda = {"unavailable": repr(exc)}

    line = {"metric": METRIC, "value": value, "unit": "segments/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f64 index/gating + f32 LUT accumulation (signals f32, pixel sums f64)", "data": "synthetic",
            "config": {"workload": WORKLOAD, "segments": S, "events": int(len(events)), "batches": int(len(unit_sizes)),
                       "batch_segments_min_mean_max": [int(unit_sizes.min()), float(unit_sizes.mean()), int(unit_sizes.max())],
                       "packets": n_packets, "data_packets": n_data, "hits": st["n_hits"] if world == 1 else None,
                       "rng": "cloud (one sample cloud per segment x pixel); per-batch states = create_xoroshiro128p_states(seed = rand_seed + batch number)",
                       "l2": "per-step working set (152 MB of records, ~1 GB of sparse waveforms per batch in flight, %.2f GB of packets + truth rows) >> 126 MB L2; "
                             "fresh device copy of the records each step" % (d2h / 1e9),
                       "pipeline": "%d batches in flight per GPU (front-end stage of one batch under the current stage of the next)" % depth,
                       "parallelism": ("%d ranks, batches assigned longest-first, no collective in the chain; NCCL send/recv of the packets + truth rows to rank 0, "
                                       "file order restored on the device" % world) if world > 1 else "single GPU"},
            "e2e": {"value": e2e_value, "unit": "segments/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": ms_e2e / args.steps,
                    "what": "SpillRunner.simulate(host records) -> host packets, mc_packets_assn rows and updated records (what fee.export_to_hdf5 "
                            "and the segments dataset receive)"},
            "gpu_launches": int(launches), "per_rank": per_rank, "clocks": clocks, "roofline": roofline, "roofline_by_kernel": by_kernel,
            "cpu_baseline": cpu, "reference_numba_cuda": ref_cuda, "module0_1e4": module0_block, "light": light_block, "kernels": kernels}
    print(json.dumps(line), flush=True)


def bench_module0(lib, ll, lchain, lc, peak, sm_mhz, steps):
    """BASELINE configs[1]: module0, 1e4 synthetic cosmic-muon segments per batch, quench -> digitize (no packets), two batches in
    flight; plus the per-kernel rooflines on that batch shape."""
    import torch
    from larndsim_b200 import synth
    mod = lc.load_snapshot("module0")
    tracks = synth.cosmic_segments(10000, mod.detector, seed=12345)
    response = synth.response_lut(mod.detector)
    S = len(tracks)
    snap = lc.snapshot()
    A, K, Tt = int(snap.max_adc_values), int(snap.max_tracks_per_pixel), int(snap.n_time_ticks)
    raw = torch.from_numpy(tracks.view(np.uint8).reshape(-1).copy()).pin_memory()
    warm = 3
    copies = [ll.DeviceRecords(dtype=tracks.dtype, n=S, buf=raw.cuda()) for _ in range(warm + steps)]
    pipe = lchain.Pipeline(tracks.dtype, response, depth=2, rng_mode="cloud")
    results = []

    def step(i):
        if pipe.full():
            results.append(pipe.collect())
        pipe.submit(copies[i], rng_seed=RAND_SEED + i)
    for i in range(warm):
        step(i)
    results += pipe.drain()
    torch.cuda.synchronize()
    results.clear()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        step(warm + i)
    results += pipe.drain()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    pipe.close()
    # per-kernel, one batch at a time
    ch = lchain.Chain(tracks.dtype, response, rng_mode="cloud", stage_timing=True)
    fresh = [ll.DeviceRecords(dtype=tracks.dtype, n=S, buf=raw.cuda()) for _ in range(steps + 1)]
    r1 = ch.run(fresh[0], rng_seed=1)
    acc = {"n_fma": 0, "n_samples": 0, "pair_ticks": 0, "pixel_ticks": 0, "U": 0, "A": A, "K": K}
    stage_acc = {}

    def passes():
        for i in range(steps):
            r = ch.run(fresh[1 + i], rng_seed=1)
            acc["n_fma"] += r.n_fma; acc["n_samples"] += r.n_samples; acc["pair_ticks"] += r.n_pairs * r.n_ticks
            acc["pixel_ticks"] += r.n_unique_pixels * Tt; acc["U"] += r.n_unique_pixels
            for k, v in r.stage_ms.items():
                stage_acc[k] = stage_acc.get(k, 0.0) + v
    prof = profile_session(lib, passes)
    ch.close()
    by_kernel = kernel_rooflines(prof, steps, acc, peak, sm_mhz)
    # the deterministic tracks_current kernel (detsim.py:351-453; not called by the CLI): bounded sample of the same batch, inputs
    # made with the drop-in kernels (quench, drift, max_pixels, get_pixels, time_intervals)
    tc = None
    try:
        from math import ceil
        from larndsim_b200 import detsim, quenching, drifting, pixels_from_track as pft
        n_tc = 128
        sub = tracks[:n_tc].copy()
        quenching.quench[1, 128](sub, int(snap.mode_birks))
        drifting.drift[1, 128](sub)
        radius = ceil(max(sub["tran_diff"]) * 5 / mod.detector.PIXEL_PITCH)
        mp = np.array([0])
        pft.max_pixels[1, 128](sub, mp)
        P_tc = int((2 * radius + 1) * mp[0] + (1 + 2 * radius) * radius * 2)
        d_sub = ll.DeviceRecords(host=sub)
        act = torch.full((n_tc, int(mp[0])), -1, dtype=torch.int32, device="cuda")
        neigh = torch.full((n_tc, P_tc), -1, dtype=torch.int32, device="cuda")
        nrad = torch.full((n_tc, P_tc), -1, dtype=torch.int32, device="cuda")
        npl = torch.zeros(n_tc, dtype=torch.float64, device="cuda")
        pft.get_pixels[1, 128](d_sub, act, neigh, nrad, npl, radius)
        ml = torch.zeros(1, dtype=torch.int64, device="cuda")
        ts = torch.empty(n_tc, dtype=torch.float64, device="cuda")
        detsim.time_intervals[1, 128](ts, ml, d_sub)
        T_tc = int(ml.item())
        sig_tc = torch.zeros((n_tc, P_tc, T_tc), dtype=torch.float32, device="cuda")
        resp_d = torch.from_numpy(response).cuda()
        detsim.tracks_current[1, 1](sig_tc, neigh, d_sub, resp_d)
        torch.cuda.synchronize()
        t0e, t1e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        sig_tc.zero_()
        t0e.record()
        detsim.tracks_current[1, 1](sig_tc, neigh, d_sub, resp_d)
        t1e.record()
        torch.cuda.synchronize()
        tc_ms = t0e.elapsed_time(t1e)
        live_pairs = int((sig_tc != 0).any(dim=2).sum().item())
        live_ticks = int((sig_tc != 0).sum().item())
        sp = int(mod.detector.SAMPLED_POINTS)
        flops = 110.0 * live_pairs * sp ** 3 + 2.0 * sp ** 3 * live_ticks          # 110 N_rho + 2 N_rho T_act, z_steps >= SAMPLED_POINTS
        fp32_peak = FP32_LANES * 2 * sm_mhz * 1e6
        tc = {"segments": n_tc, "ms": tc_ms, "segments_per_s": n_tc / (tc_ms * 1e-3), "live_pairs": live_pairs, "sampled_points": sp,
              "algorithmic_flops_lower_bound": flops, "achieved_tflops_lower_bound": flops / (tc_ms * 1e-3) / 1e12,
              "frac_of_fp32_peak_lower_bound": flops / (tc_ms * 1e-3) / fp32_peak,
              "note": "rho (erfc, exp, log) and the table products are float64 like the reference: its own ceiling is the FP64 / transcendental "
                      "rate, not FP32 FFMA"}
    except Exception as exc:
        tc = {"error": repr(exc)}
    return {"tracks_current": tc, "workload": "module0 config, synthetic cosmic-muon segments (1e4 segments per batch), quench->digitize, 2 batches in flight, noise on",
            "segments_per_s": S / (ms * 1e-3), "ms_per_batch": ms, "unique_pixels": r1.n_unique_pixels, "ticks": r1.n_ticks, "hits": r1.n_hits,
            "n_fma_per_batch": acc["n_fma"] / steps, "mc_sample_points_per_batch": acc["n_samples"] / steps,
            "stage_ms_per_batch": {k: v / steps for k, v in stage_acc.items()},
            "kernel_ms_per_batch": {k: v[1] / steps for k, v in sorted(prof.items(), key=lambda kv: -kv[1][1])[:12]},
            "roofline_by_kernel": by_kernel}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus > 1 and world == 1 and args.impl == "ours":
        # convenience: re-launch under torchrun, one rank per GPU
        port = 29500 + (os.getpid() % 500)
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.abspath(__file__), "--gpus", str(args.gpus),
               "--steps", str(args.steps), "--warmup", str(args.warmup)]
        sys.exit(subprocess.call(cmd))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        import __graft_entry__ as ge
        if not os.path.exists(ge.LIB):
            raise SystemExit("CUDA extension not built: run python __graft_entry__.py")
        run_ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
