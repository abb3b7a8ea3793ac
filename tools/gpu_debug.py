import sys, os, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import bench
from larndsim_b200 import _launch as ll, chain as lchain, consts as lc
mod, tracks, response = bench.make_batch(12345)
S = len(tracks)
raw = torch.from_numpy(tracks.view(np.uint8).reshape(-1).copy()).pin_memory()
N = 14
depth = int(os.environ.get("DEPTH", "3"))
pipe = lchain.Pipeline(tracks.dtype, response, depth=depth)
side = torch.cuda.Stream(); torch.cuda.set_stream(side)
def loop(n, show):
    devs = [ll.DeviceRecords(dtype=tracks.dtype, n=S, buf=raw.cuda()) for _ in range(n)]
    torch.cuda.synchronize()
    t0 = time.perf_counter(); tls = []
    for i in range(n):
        if pipe.full(): tls.append(pipe.collect().timeline)
        pipe.submit(devs[i], rng_seed=1)
    while pipe._inflight: tls.append(pipe.collect().timeline)
    torch.cuda.synchronize()
    print("ms/step", 1e3 * (time.perf_counter() - t0) / n)
    if show:
        base = tls[0][0]
        for t in tls: print(" ".join("%7.2f" % (x - base) for x in t), "| front %.2f mcwait %.2f mc %.2f feewait %.2f fee %.2f" % (t[1]-t[0], t[2]-t[1], t[3]-t[2], t[4]-t[3], t[5]-t[4]))
loop(6, False)
loop(N, True)
