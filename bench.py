#!/usr/bin/env python
"""Benchmark of the charge-readout chain (BASELINE.json metric: segments/s quench -> ADC).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A step = one pass of the hot path (quench -> drift -> get_pixels -> tracks_current_mc -> unique / index
maps -> sum_pixel_signals -> get_adc_values -> digitize) over one batch of synthetic segments.  Workload at
N=1: BASELINE.json configs[1] -- module0 geometry, 1e4 synthetic cosmic-muon segments, noise on.  For N>1
(torchrun, one rank per GPU) every rank processes its own batch of the same size (different muons:
independent (event, module) units), no collective inside the chain; the per-step hit
packets are gathered to rank 0 over NCCL.  ``value`` = segments of all ranks / max-over-ranks device time.

``--impl reference`` times the CPU implementation of the same path: the reference's kernels are Numba
Python and cannot travel to the GPU box, so this arm runs the C restatement pinned to them
(oracle/larnd_oracle.c, OpenMP on all host cores) on a bounded sample of the same batch.
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

CONFIG = "module0"
N_SEGMENTS = 10000
METRIC = "segments/s quench->ADC (module0, 1e4 synthetic cosmic-muon segments per batch)"


def make_batch(seed):
    from larndsim_b200 import consts as lc, synth
    mod = lc.load_snapshot(CONFIG)
    tracks = synth.cosmic_segments(N_SEGMENTS, mod.detector, seed=seed)
    response = synth.response_lut(mod.detector)
    return mod, tracks, response


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


def cpu_chain(tracks, response, n_sample, rng_seed=1):
    """The CPU restatement of the chain on the first n_sample segments; returns seconds."""
    import helpers as h
    sub = tracks[:n_sample].copy()
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


def omp_threads():
    n = os.environ.get("OMP_NUM_THREADS")
    return int(n) if n else (os.cpu_count() or 1)


def run_reference(args, rank, world):
    if rank != 0:
        return
    mod, tracks, response = make_batch(12345)
    n_sample = 2048
    for _ in range(args.warmup):
        cpu_chain(tracks, response, 32)
    times = []
    for _ in range(args.steps):
        dt, hits = cpu_chain(tracks, response, n_sample)
        times.append(dt)
    total = sum(times)
    value = n_sample * args.steps / total
    cores = omp_threads()
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "segments/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * total / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "module0 config, synthetic cosmic-muon segments (1e4 segments), charge readout quench->ADC",
                       "sample": "first %d of the 1e4-segment batch per step" % n_sample},
            "cpu_baseline": {"value": value, "unit": "segments/s", "cores": cores, "kind": "port",
                             "sample": "first %d segments of the batch, CPU restatement pinned to the reference's golden vectors "
                                       "(the reference itself is Numba Python and does not exist on the GPU box)" % n_sample},
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


def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    assert torch.cuda.is_available(), "bench.py needs a GPU (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")       # keep NCCL's banner off stdout: rank 0 prints ONE JSON line
        if os.environ.get("LSB_BENCH_BACKEND") == "gloo":       # diagnostic: control plane without NCCL (no hit gather)
            os.environ["LSB_NO_GATHER"] = "1"
            dist.init_process_group("gloo")
        else:
            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from larndsim_b200 import _launch as ll, chain as lchain, consts as lc
    lib = ll.lib()
    lib.lsb_profile_end.restype = C.c_int64
    # weak scaling: every rank simulates a batch of the same size AND the same content (independent detector modules seeing
    # identical muons), so the per-rank work is exactly fixed as N grows and max-over-ranks is not a lottery over batches
    mod, tracks, response = make_batch(12345)
    S = len(tracks)
    itemsize = tracks.dtype.itemsize
    ch = lchain.Chain(tracks.dtype, response, rng_mode="cloud", stage_timing=True)
    raw = torch.from_numpy(tracks.view(np.uint8).reshape(-1).copy())
    pinned_in = raw.pin_memory()
    n_total = args.warmup + args.steps
    dev_copies = [ll.DeviceRecords(dtype=tracks.dtype, n=S, buf=pinned_in.cuda()) for _ in range(n_total)]
    A, K = int(lc.snapshot().max_adc_values), int(lc.snapshot().max_tracks_per_pixel)

    from larndsim_b200 import dist as ldist
    gatherer = [None]

    def gather_packets(res):
        """hit table (pixel id, ADC codes, timestamps) of this batch -> rank 0 (north_star: NCCL only here);
        one fixed-size NCCL gather per batch, no host synchronisation."""
        if world == 1 or os.environ.get("LSB_NO_GATHER"):
            return
        if gatherer[0] is None:
            from larndsim_b200 import packets as lp
            ped = lp.ReadoutTables.from_consts()._c.adc_pedestal              # digitize(0): a hit is an ADC code above it
            gatherer[0] = ldist.HitTableGather(2 * int(res.n_unique_pixels) + 4096, ped, "cuda")
        gatherer[0].gather(res.unique_pix, res.adc_digit, res.adc_ticks_list)

    def sync():
        if world > 1 and not os.environ.get("LSB_BENCH_NO_BARRIER"):       # (diagnostic switch)
            dist.barrier()
        torch.cuda.synchronize()

    side = torch.cuda.Stream()
    torch.cuda.synchronize()
    torch.cuda.set_stream(side)          # everything below is timed with events on this (non-blocking) stream
    # ---------------- value: device-resident inputs, three batches in flight ----------------
    # Batches are independent, so the timed loop keeps two of them in flight (chain.Pipeline): the FEE stage of
    # batch i (latency-bound, high-priority stream) runs under the MC stage of batch i+1 (L1-bound, low priority).
    pipe = lchain.Pipeline(tracks.dtype, response, depth=2, rng_mode="cloud")
    results = []

    def collect(r):
        if r is not None:
            results.append(r)
            gather_packets(r)

    host_t = {"collect": 0.0, "submit": 0.0, "n": 0}

    def step(batch):
        w0 = time.perf_counter()
        if pipe.full():
            collect(pipe.collect())          # consume the oldest result before its chain is reused
        w1 = time.perf_counter()
        pipe.submit(batch, rng_seed=1)
        host_t["collect"] += w1 - w0; host_t["submit"] += time.perf_counter() - w1; host_t["n"] += 1

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    for i in range(args.warmup):
        step(dev_copies[i])
    while pipe._inflight:
        collect(pipe.collect())
    sync()
    results.clear()
    launches0 = lib.lsb_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host_t.update(collect=0.0, submit=0.0, n=0)
    sampler.begin()
    e0.record()
    for i in range(args.steps):
        step(dev_copies[args.warmup + i])
    while pipe._inflight:
        collect(pipe.collect())
    if gatherer[0] is not None:
        gatherer[0].flush()                  # every hit table has reached rank 0 inside the timed region
    e1.record()
    sync()
    sampler.end()
    launches = lib.lsb_launch_count() - launches0
    assert len(results) == args.steps
    ms = e0.elapsed_time(e1)
    t = torch.tensor([ms], device="cuda", dtype=torch.float64)
    per_rank = [{"ms_per_step": ms / args.steps, "mc_sample_points": int(results[-1].n_samples)}]
    if world > 1:
        # every rank simulates different muons (independent units): batches differ in their number of sample points
        rdev = "cpu" if dist.get_backend() == "gloo" else "cuda"
        t = t.to(rdev)
        allv = [torch.zeros(2, device=rdev, dtype=torch.float64) for _ in range(world)]
        dist.all_gather(allv, torch.tensor([ms / args.steps, float(results[-1].n_samples)], device=rdev, dtype=torch.float64))
        per_rank = [{"ms_per_step": float(v[0].item()), "mc_sample_points": int(v[1].item())} for v in allv]
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_max = float(t.item())
    value = world * S * args.steps / (ms_max * 1e-3)
    res = results[-1]
    U, T, P_ = res.n_unique_pixels, res.n_ticks, res.max_neighbors
    n_samples = res.n_samples
    n_hits = res.n_hits

    # ---------------- per-kernel device times: the same K steps, one batch at a time ----------------
    # (kernels of overlapping batches would stretch each other's event intervals)
    fresh = [ll.DeviceRecords(dtype=tracks.dtype, n=S, buf=pinned_in.cuda()) for _ in range(args.steps + 1)]
    ch.run(fresh[0], rng_seed=1)
    sync()
    lib.lsb_profile_begin(ll.stream())
    stage_acc = {}
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for i in range(args.steps):
        r1 = ch.run(fresh[1 + i], rng_seed=1)
        for k, v in r1.stage_ms.items():
            stage_acc[k] = stage_acc.get(k, 0.0) + v
    s1.record()
    sync()
    buf = C.create_string_buffer(1 << 16)
    lib.lsb_profile_end(buf, C.c_int64(len(buf)))
    prof = parse_profile(buf.value.decode())
    ms_serial = s0.elapsed_time(s1) / args.steps

    # ---------------- e2e: host buffers through the public chain call (pipelined the same way) ----------------
    ucap = int(U * 1.2) + 1024
    outs = [(torch.empty(ucap, dtype=torch.int32).pin_memory(), torch.empty((ucap, A), dtype=torch.float64).pin_memory(),
             torch.empty((ucap, A), dtype=torch.float64).pin_memory()) for _ in range(3)]
    host_batches = [raw.clone().pin_memory() for _ in range(3 + args.steps)]
    for i in range(3):
        if pipe.full():
            pipe.collect()
        pipe.submit_host(host_batches[i], *outs[i % 3], rng_seed=1)
    pipe.drain()
    sync()
    sampler.begin()
    e0.record()
    for i in range(args.steps):
        if pipe.full():
            pipe.collect()
        pipe.submit_host(host_batches[3 + i], *outs[i % 3], rng_seed=1)
    r2 = pipe.drain()[-1]
    e1.record()
    sync()
    sampler.end()
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e = e0.elapsed_time(e1)
    t = torch.tensor([ms_e2e], device="cpu" if (world > 1 and dist.get_backend() == "gloo") else "cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * S * args.steps / (float(t.item()) * 1e-3)
    h2d = S * itemsize
    d2h = S * itemsize + r2.n_unique_pixels * 4 + 2 * r2.n_unique_pixels * A * 8

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return
    # ---------------- next stage: hits of one batch -> LArPix packets + truth rows ----------------
    packets_block = None
    try:
        from larndsim_b200 import packets as lp
        tables = lp.ReadoutTables.from_consts()
        ev = torch.zeros((U, A), dtype=torch.int64, device="cuda")
        tpm = res.track_pixel_map
        traj = torch.where(tpm >= 0, tpm // 7, tpm)
        t0s = np.array([1000.0])
        lp.export_packets(tables, ev, res.adc_digit, res.adc_ticks_list, res.unique_pix, res.current_fractions, tpm, traj, t0s)
        torch.cuda.synchronize()
        l0 = lib.lsb_launch_count()
        w0 = time.perf_counter()
        for _ in range(args.steps):
            pk, ds = lp.export_packets(tables, ev, res.adc_digit, res.adc_ticks_list, res.unique_pix, res.current_fractions, tpm, traj, t0s)
        torch.cuda.synchronize()
        pk_ms = (time.perf_counter() - w0) * 1e3 / args.steps
        alg_bytes = 8.0 * U * A * 3 + 8.0 * n_hits * K * 3 + len(pk) * (32 + 8 * (1 + 4 * tables.association_count))
        packets_block = {"ms_per_batch": pk_ms, "packets_per_batch": int(len(pk)), "data_packets": int((pk["packet_type"] == 0).sum()),
                         "packets_per_s": len(pk) / (pk_ms * 1e-3), "launches_per_batch": (lib.lsb_launch_count() - l0) / args.steps,
                         "timed": "wall clock around the public call (device inputs, results copied to host arrays)",
                         "algorithmic_bytes": alg_bytes, "note": "latency-bound at this size (10 launches, 2 host syncs, 2 D2H copies); not part of `value`"}
        if not os.environ.get("LSB_BENCH_NO_CPU"):
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import packets_oracle as po
            import packets_util as pu                     # readout tables as plain containers (tests/)
            z = pu.load("module0")
            nsub = min(U, 2000)
            args_cpu = [x[:nsub].cpu().numpy() for x in (ev, res.adc_digit, res.adc_ticks_list, res.unique_pix, res.current_fractions, tpm, traj)]
            w0 = time.perf_counter()
            opk, _ = po.export_packets(pu.tables_from_npz(z), *args_cpu, t0s)
            cpu_s = time.perf_counter() - w0
            packets_block["cpu_python_restatement"] = {"packets_per_s": len(opk) / cpu_s, "sample": "first %d pixels, %.2f s, 1 core" % (nsub, cpu_s)}
    except Exception as exc:                                   # the packet stage is an extra: never lose the headline line
        packets_block = {"error": repr(exc)}
    # ---------------- next stage: light triggers + digitisation of one module's waveforms ----------------
    light_block = None
    try:
        from larndsim_b200 import light_sim
        import light_trigger_util as ltu
        C_l = ltu.consts_from_npz(ltu.load("module0"))
        sig_l, op_l, tid_l, tph_l = ltu.case_inputs("module0", C_l["OP_CHANNEL_PER_TRIG"], C_l["N_OP_CHANNEL"])
        thr_l = ltu.thresholds(C_l, op_l)
        sig_d = torch.from_numpy(sig_l).cuda(); tid_d = torch.from_numpy(tid_l).cuda(); tph_d = torch.from_numpy(tph_l).cuda()
        zero_noise = np.zeros((C_l["N_OP_CHANNEL"], 33))
        ns_l = 256

        def light_once():
            trig, chans, _ = light_sim.get_triggers(sig_d, thr_l, op_l, 0)
            return light_sim.sim_triggers((1, 1, 1), (1, 1, 64), sig_d, op_l, tid_d, tph_d, trig, chans, ns_l, zero_noise), trig
        light_once()
        torch.cuda.synchronize()
        w0 = time.perf_counter()
        for _ in range(args.steps):
            (dg, _, _), trig_l = light_once()
        torch.cuda.synchronize()
        l_ms = (time.perf_counter() - w0) * 1e3 / args.steps
        light_block = {"ms_per_call": l_ms, "channels": int(sig_l.shape[0]), "ticks": int(sig_l.shape[1]), "triggers": int(len(trig_l)),
                       "digitised_samples": int(dg.numel()), "workload": "tests/light_trigger_util.py case module0 (96 channels x 9000 ticks, 2 truth slots)",
                       "timed": "wall clock around get_triggers + sim_triggers (device inputs)",
                       "algorithmic_bytes": 4.0 * sig_l.size + 8.0 * dg.numel() * (1 + 2 * tid_l.shape[2])}
        if not os.environ.get("LSB_BENCH_NO_CPU"):
            sys.path.insert(0, os.path.join(ROOT, "oracle"))
            import light_trigger_oracle as lo
            w0 = time.perf_counter()
            t_o, c_o, _ = lo.get_triggers(sig_l, thr_l, op_l, 0, C_l)
            lo.sim_triggers(sig_l, op_l, tid_l, tph_l, t_o[:1], c_o[:1, :8], ns_l, C_l)
            cpu_s = time.perf_counter() - w0
            light_block["cpu_python_restatement"] = {"seconds": cpu_s, "sample": "trigger search on all channels + digitisation of 1 trigger x 8 channels (of %d x %d), 1 core" % (len(t_o), c_o.shape[1])}
    except Exception as exc:
        light_block = {"error": repr(exc)}
    # ---------------- BASELINE metric, second half: the deterministic tracks_current kernel (detsim.py:351-453) ----------------
    # (not called by the CLI; a bounded sample of the same batch: SAMPLED_POINTS^2 x z_steps rho evaluations per pair and tick slab)
    tc_block = None
    try:
        import helpers as hh
        from larndsim_b200 import detsim, quenching, drifting, pixels_from_track as pft
        n_tc = 128
        sub = tracks[:n_tc].copy()
        d_sub = ll.DeviceRecords(host=sub)
        quenching.quench[1, 1](d_sub, int(lc.snapshot().mode_birks))
        drifting.drift[1, 1](d_sub)
        fr = hh.oracle_front(sub.copy(), hh.Oracle(lc.snapshot()))         # pixel lists / tick count of the sample (host)
        neigh = torch.from_numpy(fr["neigh"]).cuda()
        sig_tc = torch.zeros((n_tc, fr["P"], fr["T"]), dtype=torch.float32, device="cuda")
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
        n_rho = float(live_pairs) * sp * sp * sp                              # z_steps >= SAMPLED_POINTS: a lower bound
        flops = 110.0 * n_rho + 2.0 * sp * sp * sp * live_ticks                # 110 N_rho + 2 N_rho T_act
        sm_clk_tc = 1965.0
        tc_block = {"segments": n_tc, "ms": tc_ms, "segments_per_s": n_tc / (tc_ms * 1e-3), "live_pairs": live_pairs,
                    "flops_lower_bound": flops, "achieved_tflops_lower_bound": flops / (tc_ms * 1e-3) / 1e12,
                    "fp32_peak_tflops": 148 * 128 * 2 * sm_clk_tc * 1e6 / 1e12, "frac_of_fp32_peak_lower_bound": flops / (tc_ms * 1e-3) / (148 * 128 * 2 * sm_clk_tc * 1e6),
                    "note": "the kernel evaluates rho (erf, exp, log) and the table products in float64 like the reference, so its own ceiling is the "
                            "FP64 / transcendental rate, not FP32 FFMA; flops count the algorithmic formulation with z_steps = SAMPLED_POINTS"}
    except Exception as exc:
        tc_block = {"error": repr(exc)}
    # ---------------- roofline of the dominant kernel ----------------
    peak, peak_src = measured_peaks()
    Tt = int(lc.snapshot().n_time_ticks)
    kern = [(k, v) for k, v in prof.items() if not k.startswith("(")]
    total_kernel_ms = sum(v[1] for _, v in kern)
    top_name, (top_cnt, top_ms) = max(kern, key=lambda kv: kv[1][1])
    per_launch_ms = top_ms / max(top_cnt, 1)
    launches_per_step = top_cnt / args.steps
    # algorithmic HBM bytes per launch.  n_valid = (segment,pixel) pairs with a pixel,
    # T_act = ticks with time >= 0 actually written (about half of T for uniformly distributed drift times)
    n_valid = 0.61 * S * P_
    alg = {
        # out: the ticks of signals the samples cover (rows are stored sparsely, about half of T per valid pair); in: group
        # records (16 B, ~1 per 3 samples), sample records for the edge ticks (24 B), pair records
        "k_mc_accumulate": 4.0 * n_valid * T * 0.5 + 24.0 * n_samples + 16.0 * n_samples / 3.0 + 160.0 * S * P_,
        "k_fee_trigger": 8.0 * U * Tt + 4.0 * U * 4480 + 2 * 8.0 * U * A,
        "k_sum_pixel_signals": 4.0 * n_valid * T + 2 * 8.0 * U * Tt,
        "k_mc_sampler": 24.0 * n_samples + 28.0 * n_samples,
        "k_mc_uniforms": 24.0 * n_samples,
        "k_mc_sort": 4.0 * n_samples + 16.0 * n_samples / 3.0,
    }
    traffic = None
    ncu = {}
    tpath = os.path.join(ROOT, "profiles", "r01_traffic.json")
    if os.path.exists(tpath):
        with open(tpath) as f:
            tj = json.load(f)
        ncu = tj.get(top_name, {})
        traffic = ncu.get("dram_bytes")
    bytes_per_launch = alg.get(top_name)
    roofline = {"kernel": top_name, "bound": "hbm", "timed": "CUDA events on the launching stream, same K steps run one batch at a time",
                "share_of_kernel_time": top_ms / total_kernel_ms, "ms_per_launch": per_launch_ms, "launches_per_step": launches_per_step,
                "achieved": (bytes_per_launch / (per_launch_ms * 1e-3) / 1e9) if bytes_per_launch else None,
                "peak": peak, "peak_source": peak_src, "unit": "GB/s", "traffic": traffic,
                "traffic_source": tj.get("_source") if traffic else None}
    roofline["frac"] = roofline["achieved"] / peak if roofline["achieved"] else None
    if top_name == "k_mc_accumulate":
        # This kernel is not HBM-bound (the table is L2/L1-resident): its limiter is the L1TEX data pipe.  The HBM figures
        # above are reported because the contract asks for them; these explain the kernel.  Algorithmic work:
        # N_fma sample-tick pairs, one table word and one add each; the grouped path serves them with ~1 aligned LDG.128
        # per 2.4 pairs x 32 lanes and a count-weighted FFMA per distinct offset.
        n_fma = 32.0 * 9.19e8 * (S * P_ / 180000.0)      # sample-tick pairs per launch (ncu count of the one-load-per-pair kernel)
        sm_clk = (clocks or {}).get("sm_mhz") or 1965.0
        mc_ms = sum(prof[k][1] for k in ("k_mc_pairs", "k_mc_uniforms", "k_mc_sampler", "k_mc_sort", "k_mc_accumulate") if k in prof) / args.steps
        fp32_peak = 148 * 128 * 2 * sm_clk * 1e6 / 1e12
        roofline["limiter"] = {"what": "L1TEX data pipe (table gather, table L2-resident)",
                               "sample_tick_pairs_per_launch": n_fma, "pairs_per_s": n_fma / (per_launch_ms * 1e-3),
                               "l1_wavefront_peak_per_s": 148 * sm_clk * 1e6,
                               "ncu_l1tex_throughput_pct": ncu.get("l1tex_throughput_pct"),
                               "ncu_l1_global_load_requests": ncu.get("l1_global_load_requests"),
                               "ncu_issue_active_pct": ncu.get("issue_active_pct")}
        # BASELINE metric, second half: tracks_current_mc against the FP32 peak, algorithmic FLOPs of SURVEY 8(d)
        flops = 2.0 * n_fma + 330.0 * n_samples
        roofline["tracks_current_mc_fp32"] = {"algorithmic_flops_per_batch": flops, "stage_ms": mc_ms,
                                              "achieved_tflops": flops / (mc_ms * 1e-3) / 1e12, "peak_tflops": fp32_peak,
                                              "frac": flops / (mc_ms * 1e-3) / 1e12 / fp32_peak,
                                              "peak_source": "148 SM x 128 lanes x 2 x SM clock (no tensor cores: nothing is a contraction)"}
    kernels = {k: {"launches_per_step": v[0] / args.steps, "ms_per_step": v[1] / args.steps} for k, v in
               sorted(prof.items(), key=lambda kv: -kv[1][1])[:12]}

    # ---------------- CPU baseline (bounded sample) ----------------
    cpu = None
    if world == 1 and not os.environ.get("LSB_BENCH_NO_CPU"):     # (the switch is for kernel A/B runs, tools/ab_variants.sh)
        n_cpu = 4096
        cpu_chain(tracks, response, 16)
        cpu_s, _ = cpu_chain(tracks, response, n_cpu)
        cpu = {"value": n_cpu / cpu_s, "unit": "segments/s", "cores": omp_threads(), "kind": "port",
               "sample": "first %d segments of the same batch, %.1f s; C/OpenMP restatement of the reference kernels (oracle/), "
                         "pinned to the reference's golden vectors" % (n_cpu, cpu_s)}

    line = {"metric": METRIC, "value": value, "unit": "segments/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_max / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64 index/gating + f32 LUT accumulation (signals f32, pixel sums f64)", "data": "synthetic",
            "config": {"workload": "module0 config, synthetic cosmic-muon segments (1e4 segments), charge readout quench->ADC, noise on",
                       "segments_per_batch": S, "pixels_per_segment_row": P_, "unique_pixels": U, "ticks": T, "hits": n_hits,
                       "mc_sample_points": n_samples, "rng": "cloud (one sample cloud per segment x pixel)",
                       "l2": "per-step working set (signals %.2f GB dense-equivalent, per-segment pixel waveforms %.2f GB) >> 126 MB L2; fresh input copy each step"
                             % (4.0 * S * P_ * T / 1e9, 8.0 * U * Tt * K / 1e9),
                       "pipeline": "2 batches in flight per GPU (FEE stage of batch i under the MC stage of batch i+1)",
                       "parallelism": "1 batch stream per rank (identical batch on every rank), no collective in the chain; NCCL gather of the compacted hit packets to rank 0" if world > 1 else "single GPU"},
            "e2e": {"value": e2e_value, "unit": "segments/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                    "ms_per_step": float(t.item()) / args.steps},
            "gpu_launches": int(launches), "per_rank": per_rank, "host_ms_per_step": {"wait_for_oldest_batch": 1e3 * host_t["collect"] / max(host_t["n"], 1),
                                                              "submit_next_batch": 1e3 * host_t["submit"] / max(host_t["n"], 1)},
            "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "packets": packets_block, "light_triggers": light_block, "tracks_current": tc_block,
            "ms_per_step_unpipelined": ms_serial,
            "stage_ms_per_step": {k: v / args.steps for k, v in stage_acc.items()}, "kernels": kernels}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
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
