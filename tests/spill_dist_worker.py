"""torchrun worker: the partitioned spill (N ranks, NCCL) must give rank 0 exactly the bytes a single rank produces.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 tests/spill_dist_worker.py
"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from larndsim_b200 import consts as lc, synth, spill  # noqa: E402
import helpers  # noqa: E402


def main():
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    import datetime
    dist.init_process_group("nccl", device_id=torch.device("cuda", local), timeout=datetime.timedelta(seconds=120))
    rank, world = dist.get_rank(), dist.get_world_size()
    ok = True
    for config, n, nev in (("ndlar", 30000, 3), ("2x2", 8000, 2)):
        mod = lc.load_snapshot(config)
        resp = synth.response_lut(mod.detector)
        tracks = synth.beam_spill_segments(n, mod.detector, seed=99, n_events=nev)
        tracks["segment_id"] = np.arange(len(tracks)); tracks["file_traj_id"] = tracks["traj_id"]
        runner = spill.SpillRunner(tracks.dtype, resp, depth=3)
        out = runner.simulate(tracks.copy(), rand_seed=5, return_tracks=True)
        if rank == 0:                                   # the host arrays are views of staging buffers the next call reuses
            out.packets, out.packets_mc_ds, out.tracks = out.packets.copy(), out.packets_mc_ds.copy(), out.tracks.copy()
        again = runner.simulate(tracks.copy(), rand_seed=5)
        if rank == 0:
            again.packets = again.packets.copy()
        # result left on the device: the NCCL exchange to rank 0 + device-side file order (host output above went through the
        # host table shared by the ranks)
        dev = runner.simulate(tracks.copy(), rand_seed=5, host_output=False)
        if rank == 0:
            single = spill.SpillRunner(tracks.dtype, resp, depth=2, single_rank=True)
            ref = single.simulate(tracks.copy(), rand_seed=5, return_tracks=True)
            same = (out.packets.tobytes() == ref.packets.tobytes() and out.packets_mc_ds.tobytes() == ref.packets_mc_ds.tobytes()
                    and again.packets.tobytes() == ref.packets.tobytes() and np.array_equal(out.unit_packets, ref.unit_packets)
                    and dev.packets.cpu().numpy().tobytes() == ref.packets.tobytes()
                    and dev.packets_mc_ds.cpu().numpy().tobytes() == ref.packets_mc_ds.tobytes()
                    and helpers.records_equal(out.tracks, ref.tracks))          # field-wise: padding bytes are not data
            print("%s: %d ranks, %d units, %d packets, rank-0 output == single-rank output: %s" % (
                config, world, len(out.unit_sizes), out.n_packets, same), flush=True)
            if not same:
                plan = spill.assign_units(out.unit_sizes, world)
                owner = {u: r for r, lst in enumerate(plan) for u in lst}
                print("  unit_packets equal:", np.array_equal(out.unit_packets, ref.unit_packets), " tracks equal:", helpers.records_equal(out.tracks, ref.tracks),
                      " again==out:", again.packets.tobytes() == out.packets.tobytes(), " n:", len(out.packets), len(ref.packets))
                if len(out.packets) == len(ref.packets):
                    a, b = out.packets.view(np.uint8).reshape(len(out.packets), -1), ref.packets.view(np.uint8).reshape(len(ref.packets), -1)
                    bad = np.nonzero((a != b).any(axis=1))[0]
                    ra, rb = out.packets_mc_ds.view(np.uint8).reshape(len(out.packets), -1), ref.packets_mc_ds.view(np.uint8).reshape(len(ref.packets), -1)
                    badr = np.nonzero((ra != rb).any(axis=1))[0]
                    print("  differing packets:", len(bad), bad[:10], " differing rows:", len(badr), badr[:10])
                    # which unit holds the first differing packet
                    evp = 0
                    starts = []
                    pos = 0
                    nB = len(out.unit_sizes) // len(np.unique(tracks["event_id"]))
                    for u, n in enumerate(out.unit_packets):
                        if u % nB == 0:
                            pos += 2          # timestamp + trigger packet of a new event (no sync packets due in this test)
                        starts.append(pos); pos += int(n)
                    starts = np.array(starts)
                    for idx in list(bad[:3]) + list(badr[:3]):
                        u = int(np.searchsorted(starts, idx, side="right") - 1)
                        print("   index", idx, "unit", u, "owner", owner.get(u), "unit size", out.unit_sizes[u], "unit packets", out.unit_packets[u], "offset in unit", idx - starts[u])
                    if len(bad):
                        print("   out:", out.packets[bad[0]], "\n   ref:", ref.packets[bad[0]])
                    if len(badr):
                        print("   out row:", out.packets_mc_ds[badr[0]], "\n   ref row:", ref.packets_mc_ds[badr[0]])
            ok = ok and same and out.n_packets > 1000
            single.close()
        else:
            assert out.packets is None
        runner.close()
        dist.barrier()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
