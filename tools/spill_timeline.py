"""Per-unit timeline of one spill through the runner (LSB_SPILL_TIMELINE): where the GPU time between the MC stages goes.
    python tools/spill_timeline.py [depth]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
path = os.path.join(ROOT, "gpurun_out", "spill_timeline.txt")
os.makedirs(os.path.dirname(path), exist_ok=True)
if os.path.exists(path):
    os.remove(path)
import torch, bench
from larndsim_b200 import spill, _launch as ll
depth = int(sys.argv[1]) if len(sys.argv) > 1 else 3
mod, tracks, resp = bench.make_spill()
ev = np.unique(tracks["event_id"])
r = spill.SpillRunner(tracks.dtype, resp, depth=depth)
for i in range(2):
    r.simulate(tracks, events=ev, rand_seed=1, host_output=False)
torch.cuda.synchronize()
os.environ["LSB_SPILL_TIMELINE"] = path
import time
t0 = time.perf_counter()
r.simulate(tracks, events=ev, rand_seed=1, host_output=False)
torch.cuda.synchronize()
print("wall ms", (time.perf_counter() - t0) * 1e3)
rows = np.array([[float(x) for x in l.split()] for l in open(path) if not l.startswith("#")])
fb, fe, mb, me, eb, dn = rows[:, 2], rows[:, 3], rows[:, 4], rows[:, 5], rows[:, 6], rows[:, 7]
o = np.argsort(mb)
mb, me, fb, fe, eb, dn = mb[o], me[o], fb[o], fe[o], eb[o], dn[o]
print("units", len(rows), "span ms", dn.max() - fb.min())
print("sum MC ms", (me - mb).sum(), " sum front", (fe - fb).sum(), " sum FEE+export", (dn - eb).sum())
gaps = mb[1:] - me[:-1]
print("MC-to-MC gaps: sum %.1f mean %.3f max %.3f" % (gaps.sum(), gaps.mean(), gaps.max()))
print("MC start - front end (this unit): mean %.3f" % (mb - fe).mean())
print("FEE begin - MC end: mean %.3f ; done - FEE begin: mean %.3f" % ((eb - me).mean(), (dn - eb).mean()))
for i in range(6, 12):
    print("unit %d: front %.2f-%.2f MC %.2f-%.2f FEE %.2f-%.2f" % (i, fb[i] - fb[0], fe[i] - fb[0], mb[i] - fb[0], me[i] - fb[0], eb[i] - fb[0], dn[i] - fb[0]))
