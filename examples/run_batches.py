"""A whole input file through the charge path with the native batch loop (`larndsim_b200.spill.SpillRunner`): segments of a
multi-event file -> active volume cut -> quench / drift -> (event, TPC group) batches -> chain on the GPU(s) -> LArPix packets
+ `mc_packets_assn` rows on rank 0, in the reference's file order.  Single GPU, or one rank per GPU under torchrun:

    python examples/run_batches.py --config ndlar --segments 200000 --events 4
    python -m torch.distributed.run --nproc-per-node 4 --master-addr 127.0.0.1 examples/run_batches.py --config ndlar --segments 1000000

examples/reference_loop.py is the same loop written call by call with the drop-in modules (what cli/simulate_pixels.py runs
after the import switch); both give the same bytes (tests/test_gpu_spill.py).
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from larndsim_b200 import consts, synth, spill  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="2x2")
    ap.add_argument("--segments", type=int, default=20000)
    ap.add_argument("--events", type=int, default=4)
    ap.add_argument("--depth", type=int, default=3)
    a = ap.parse_args()
    import torch
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    if world > 1:
        import torch.distributed as tdist
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        tdist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0"))))
    mod = consts.load_snapshot(a.config)
    tracks = synth.beam_spill_segments(a.segments, mod.detector, seed=12345, n_events=a.events)     # every rank reads the same "file"
    tracks["segment_id"] = np.arange(len(tracks))
    tracks["file_traj_id"] = tracks["traj_id"]
    runner = spill.SpillRunner(tracks.dtype, synth.response_lut(mod.detector), depth=a.depth)
    for rep in range(2):          # first pass: device buffers grow to the largest batch; second pass: steady state
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        out = runner.simulate(tracks, rand_seed=1, return_tracks=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print("pass %d, rank %d of %d: %.3f s, %d batches / %d segments simulated here" % (
            rep, rank, world, dt, out.stats["n_units_here"], out.stats["n_segments_here"]), flush=True)
    if rank == 0:
        pk = out.packets
        kinds, counts = np.unique(pk["packet_type"], return_counts=True)
        print("segments %d  batches %d (non-empty %d)  packets %d %s  truth rows %d  -> %.0f segments/s whole loop" % (
            out.n_segments, len(out.unit_sizes), int((out.unit_sizes > 0).sum()), len(pk), dict(zip(kinds.tolist(), counts.tolist())),
            len(out.packets_mc_ds), out.n_segments / dt))
    runner.close()
    if world > 1:
        tdist.barrier()
        tdist.destroy_process_group()


if __name__ == "__main__":
    main()
