#!/bin/bash
# A/B of kernel build variants on the GPU box: tools/ab_variants.sh build/v0.so build/v1.so ...
# (each variant is copied over the product library, bench.py is run, the per-kernel times are printed)
LIB=larndsim_b200/csrc/liblarndsim_b200.so
cp $LIB /tmp/lib_orig.so
for v in "$@"; do
  cp "$v" $LIB
  LSB_BENCH_NO_CPU=1 python bench.py --steps 6 --warmup 3 > gpurun_out/ab_$(basename $v .so).json 2> gpurun_out/ab_$(basename $v .so).err
  python - "$v" <<PY
import json,sys
v=sys.argv[1]
import os
f="gpurun_out/ab_"+os.path.basename(v)[:-3]+".json"
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    k=d["kernels"]
    print(v, "ms/step %.2f unpiped %.2f |" % (d["ms_per_step"], d["ms_per_step_unpipelined"]), " ".join("%s %.3f" % (n.replace("k_", ""), x["ms_per_step"]) for n, x in list(k.items())[:8]))
except Exception as e:
    print(v, "FAILED", e)
PY
done
cp /tmp/lib_orig.so $LIB
