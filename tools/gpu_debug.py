import sys, os
sys.path.insert(0, "tests")
import numpy as np, torch
import helpers as h
for noise in (True, False):
    r = h.chain_vs_oracle(n_segments=200, config="module0", seed=17, noise=noise)
    print(noise, {k: v for k, v in r.items() if k.startswith("adc") or k.startswith("n_hits") or k.startswith("cf")})
