"""Print the per-batch stage timeline of the pipelined chain (front begin/end, MC begin/end, FEE begin, done), ms
relative to the first batch: shows how much of the FEE stage of batch i runs under the MC stage of batch i+1."""
import sys, os
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from larndsim_b200 import chain as lchain, consts as lc, synth, _launch as ll

mod = lc.load_snapshot("module0")
tracks = synth.cosmic_segments(10000, mod.detector, seed=12345)
resp = synth.response_lut(mod.detector)
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 2
pipe = lchain.Pipeline(tracks.dtype, resp, depth=depth, rng_mode="cloud")
raw = torch.from_numpy(tracks.view(np.uint8).reshape(-1).copy()).pin_memory()
n = 12
batches = [ll.DeviceRecords(dtype=tracks.dtype, n=len(tracks), buf=raw.cuda()) for _ in range(n)]
out = []
for b in batches:
    if pipe.full():
        out.append(pipe.collect())
    pipe.submit(b, rng_seed=1)
out += pipe.drain()
t0 = out[4].timeline[0]
print("batch  front0  front1     mc0     mc1    fee0    done | mc   fee  period")
prev = None
for i, r in enumerate(out[4:]):
    t = [x - t0 for x in r.timeline]
    per = (t[5] - prev) if prev is not None else float("nan")
    prev = t[5]
    print("%5d %7.2f %7.2f %7.2f %7.2f %7.2f %7.2f | %4.2f %4.2f %5.2f" % ((i,) + tuple(t) + (t[3] - t[2], t[5] - t[4], per)))
