"""One module0 batch (1e4 cosmics) and one ND-LAr (event, TPC pair) unit of the bench spill through the chain, for ncu captures
and MC diagnostics (samples per group record, share of edge ticks).   python tools/profile_units.py [module0|ndlar|both]"""
import os, sys, json
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
from larndsim_b200 import consts as lc, synth, chain as lchain, _launch as ll

which = sys.argv[1] if len(sys.argv) > 1 else "both"
reps = int(os.environ.get("REPS", 2))
out = {}
if which in ("module0", "both"):
    mod = lc.load_snapshot("module0")
    tr = synth.cosmic_segments(10000, mod.detector, seed=12345)
    ch = lchain.Chain(tr.dtype, synth.response_lut(mod.detector), stage_timing=True)
    for _ in range(reps):
        r = ch.run(ll.DeviceRecords(host=tr.copy()), rng_seed=1)
    out["module0"] = dict(S=r.n_segments, P=r.max_neighbors, U=r.n_unique_pixels, T=r.n_ticks, n_samples=r.n_samples, n_fma=r.n_fma,
                          n_pairs=r.n_pairs, n_groups=r.n_groups, n_edge=r.n_edge, n_irregular=r.n_irregular, stage_ms=r.stage_ms)
    ch.close()
if which in ("ndlar", "both"):
    mod, tracks, resp = bench.make_spill()
    if os.environ.get("THR_SCALE"):          # e.g. 1000: no pixel ever triggers (what the FEE state machine costs without hits)
        mod.detector.DISCRIMINATION_THRESHOLD *= float(os.environ["THR_SCALE"])
    sub = bench.cpu_sample(tracks, mod, 7000)
    ch = lchain.Chain(sub.dtype, resp, stage_timing=True)
    for _ in range(reps):
        r = ch.run(ll.DeviceRecords(host=sub.copy()), rng_seed=1)
    out["ndlar"] = dict(S=r.n_segments, P=r.max_neighbors, U=r.n_unique_pixels, T=r.n_ticks, n_samples=r.n_samples, n_fma=r.n_fma,
                        n_pairs=r.n_pairs, n_groups=r.n_groups, n_edge=r.n_edge, n_irregular=r.n_irregular, stage_ms=r.stage_ms)
    ch.close()
if os.environ.get("KPROF"):
    # per-kernel times of one more pass over the last unit (event pairs around every launch, lsb_profile_begin/end)
    import ctypes as C
    lib = ll.lib()
    lib.lsb_profile_end.restype = C.c_int64
    name = "ndlar" if which in ("ndlar", "both") else "module0"
    if name == "ndlar":
        ch = lchain.Chain(sub.dtype, resp); data = sub
    else:
        ch = lchain.Chain(tr.dtype, synth.response_lut(mod.detector)); data = tr
    ch.run(ll.DeviceRecords(host=data.copy()), rng_seed=1)
    torch.cuda.synchronize()
    lib.lsb_profile_begin(ll.stream())
    ch.run(ll.DeviceRecords(host=data.copy()), rng_seed=1)
    torch.cuda.synchronize()
    buf = C.create_string_buffer(1 << 16)
    lib.lsb_profile_end(buf, C.c_int64(len(buf)))
    prof = {}
    for ln in buf.value.decode().strip().split("\n"):
        f = ln.split()
        if len(f) == 3:
            prof[f[0]] = round(float(f[2]), 4)
    out[name]["kernel_ms"] = dict(sorted(prof.items(), key=lambda kv: -kv[1])[:12])
    ch.close()
for k, v in out.items():
    v["samples_per_group"] = v["n_samples"] / max(v["n_groups"], 1)
    v["edge_share_of_fma"] = v["n_edge"] / max(v["n_fma"], 1)
    v["fma_per_segment"] = v["n_fma"] / v["S"]
print(json.dumps(out, indent=1))
