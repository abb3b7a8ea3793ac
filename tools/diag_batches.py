"""Per-batch wall time and stage times of the chain over the batches of examples/run_batches.py (diagnostic)."""
import importlib.util, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
spec = importlib.util.spec_from_file_location("run_batches", os.path.join(ROOT, "examples", "run_batches.py"))
rb = importlib.util.module_from_spec(spec); spec.loader.exec_module(rb)
config, nseg, nev = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
mod = rb.consts.load_snapshot(config)
tracks = rb.synth.beam_spill_segments(nseg, mod.detector, seed=12345, n_events=nev)
keep = rb.active_volume.select_active_volume(tracks, mod.detector.TPC_BORDERS); tracks = np.ascontiguousarray(tracks[keep])
b = rb.batching.TPCBatcher(tracks, tracks, "event_id", tpc_batch_size=2, tpc_borders=mod.detector.TPC_BORDERS)
ch = rb.chain_mod.Chain(tracks.dtype, rb.synth.response_lut(mod.detector), stage_timing=len(sys.argv) > 4)
for rep in range(2):
    for k, (ev, idx) in enumerate(b.units()):
        if len(idx) == 0: continue
        sub = rb.launch.DeviceRecords(host=np.ascontiguousarray(tracks[idx]))
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = ch.run(sub, rng_seed=1 + int(ev), n_events=1)
        torch.cuda.synchronize(); t1 = time.perf_counter()
        print("rep %d batch %d S=%d U=%d T=%d P=%d samples %d wall %.1f ms | %s" % (rep, k, len(idx), res.n_unique_pixels, res.n_ticks, res.max_neighbors, res.n_samples, (t1 - t0) * 1e3,
              " ".join("%s %.2f" % (n[:10], v) for n, v in res.stage_ms.items() if v > 0.3)), flush=True)
