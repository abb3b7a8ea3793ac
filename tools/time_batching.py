import sys, time, importlib, numpy as np, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import batching_util as bu
from oracle import batching_oracle as orc
av = importlib.import_module("larndsim_b200.active_volume"); bt = importlib.import_module("larndsim_b200.util.batching")
n = 1_000_000
seg = bu.segments("ndlar", n, "f4", 99); borders = bu.borders_of("ndlar")
for rep in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    it = bt.TPCBatcher(seg, seg, "event_id", tpc_batch_size=2, tpc_borders=borders); s = it.unit_sizes
    torch.cuda.synchronize(); t1 = time.perf_counter()
    print("TPCBatcher plan 1e6 segs, %d units: %.1f ms (host records: includes the %d MB upload + %d MB order download)" % (len(s), (t1 - t0) * 1e3, seg.nbytes >> 20, n * 8 >> 20))
# device only
d = torch.from_numpy(seg.view(np.uint8).copy()).cuda()
class DR:
    def __init__(s, a, t): s.dtype = a.dtype; s.__cuda_array_interface__ = {"shape": a.shape, "typestr": "|V%d" % a.dtype.itemsize, "descr": a.dtype.descr, "data": (t.data_ptr(), False), "version": 3}
dr = DR(seg, d)
for rep in range(3):
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record(); first, idx, _ = av.classify(dr, borders, want_indices=False); e1.record(); torch.cuda.synchronize()
    print("k_active_volume 1e6 x 70 TPC: %.3f ms" % e0.elapsed_time(e1))
t0 = time.perf_counter(); sub = seg[:20000]; b = orc.tpc_batches(sub, sub, "event_id", 2, borders); t1 = time.perf_counter()
print("oracle (NumPy, reference algorithm) on 2e4 segments: %.2f s -> x50 for 1e6" % (t1 - t0))
