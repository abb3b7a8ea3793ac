import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import numpy as np, torch, helpers as h
import test_gpu_kernels as t
from larndsim_b200 import detsim
mod, tr, orc, front, resp = t._mc_setup(3, "module0")
mod.detector.SAMPLED_POINTS = 8
orc = h.Oracle()
S, P_ = front["neigh"].shape; T = front["T"]
ref = orc.tracks_current(tr, front["neigh"], T, resp)
sig = torch.zeros((S, P_, T), dtype=torch.float32, device="cuda")
detsim.tracks_current[(S, P_, (T + 63) // 64), (1, 1, 64)](sig, front["neigh"], tr, resp)
got = sig.cpu().numpy()
a=got.astype(np.float64); b=ref.astype(np.float64)
scale=np.abs(b).max(axis=-1,keepdims=True); den=np.abs(b)+1e-2*scale; den[den==0]=1
e=np.abs(a-b)/den
pe=e.max(axis=-1)
print("max", e.max(), "pairs>1e-5", (pe>1e-5).sum(), "of", (scale[...,0]>0).sum())
w=np.unravel_index(np.argmax(e), e.shape); print(w, a[w], b[w], scale[w[0],w[1]], "global peak", np.abs(b).max())
print("peaks of bad pairs", scale[...,0][pe>1e-5], "errs", pe[pe>1e-5])
# who is off: oracle or kernel, against the reference's own output (golden, 1 segment, SAMPLED_POINTS as stored)
import test_oracle_golden as tog
from larndsim_b200 import consts as lc2
g, mod = tog.load("current_fee_module0", "module0")
orc = h.Oracle()
ref = g["tc:signals"]; T = ref.shape[2]
o = orc.tracks_current(g["tracks"][:1], g["neigh"][:1], T, g["lut"])
sig = np.zeros_like(ref)
detsim.tracks_current[(1, ref.shape[1], T), (1, 1, 1)](sig, g["neigh"][:1], g["tracks"][:1], g["lut"])
print("SAMPLED_POINTS", mod.detector.SAMPLED_POINTS, "oracle vs golden", h.rel_err(o, ref), "kernel vs golden", h.rel_err(sig, ref), "kernel vs oracle", h.rel_err(sig, o))
