import sys, os, time
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import numpy as np, torch
import bench
from larndsim_b200 import _launch as ll, chain as lchain, consts as lc
mod, tracks, response = bench.make_batch(12345)
S = len(tracks)
raw = torch.from_numpy(tracks.view(np.uint8).reshape(-1).copy()).pin_memory()
N = 16
devs = [ll.DeviceRecords(dtype=tracks.dtype, n=S, buf=raw.cuda()) for _ in range(N)]
pipe = lchain.Pipeline(tracks.dtype, response, depth=int(os.environ.get("DEPTH","2")))
def loop(tag, n, record=False):
    torch.cuda.synchronize()
    t0 = time.perf_counter(); log = []
    for i in range(n):
        a = time.perf_counter()
        if pipe.full():
            pipe.collect()
        b = time.perf_counter()
        pipe.submit(devs[i], rng_seed=1)
        c = time.perf_counter()
        log.append((b - a, c - b))
    while pipe._inflight: pipe.collect()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(tag, "ms/step", 1e3 * dt / n, " collect/submit ms:", [(round(1e3*x,2), round(1e3*y,2)) for x, y in log[-4:]])
loop("warm", 4)
devs = [ll.DeviceRecords(dtype=tracks.dtype, n=S, buf=raw.cuda()) for _ in range(N)]
loop("dev-pipelined", N)
devs = [ll.DeviceRecords(dtype=tracks.dtype, n=S, buf=raw.cuda()) for _ in range(N)]
e0 = torch.cuda.Event(enable_timing=True); e0.record()
loop("dev-pipelined after legacy e0.record", N)
devs = [ll.DeviceRecords(dtype=tracks.dtype, n=S, buf=raw.cuda()) for _ in range(N)]
torch.cuda.synchronize()
with torch.cuda.stream(torch.cuda.Stream()):
    e0 = torch.cuda.Event(enable_timing=True); e0.record()
    loop("dev-pipelined on torch side stream + e0.record", N)

# host variant
A = int(lc.snapshot().max_adc_values)
outs = [(torch.empty(30000, dtype=torch.int32).pin_memory(), torch.empty((30000, A), dtype=torch.float64).pin_memory(), torch.empty((30000, A), dtype=torch.float64).pin_memory()) for _ in range(2)]
hb = [raw.clone().pin_memory() for _ in range(N)]
torch.cuda.synchronize(); t0 = time.perf_counter()
for i in range(N):
    if pipe.full(): pipe.collect()
    pipe.submit_host(hb[i], *outs[i % 2], rng_seed=1)
pipe.drain(); torch.cuda.synchronize()
print("host-pipelined ms/step", 1e3 * (time.perf_counter() - t0) / N)
