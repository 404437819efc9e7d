#!/bin/bash
# Timing ablation of the tcgen05 conv kernel (results are WRONG with any bit set; timing only):
# SRF_IGEMM_DBG bits: 1 skip A gathers, 2 skip W_k copies, 4 skip MMA issue, 8 skip epilogue stores,
# 16 skip index LDS, 32 mbarrier.arrive instead of tcgen05.commit, 64 plain arrive instead of cp.async arrive,
# 128 one arriving lane per producer warp
for dbg in ${ABLATE_SET:-0 1 2 4 8 3 15}; do
  out=$(SRF_IGEMM_DBG=$dbg timeout 200 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --layers-out /tmp/layers.json "$@" 2>&1 | tail -1)
  echo "dbg=$dbg $(python -c '
import json, collections
rows = json.load(open("/tmp/layers.json"))["layers"]
g = collections.OrderedDict()
for r in rows:
    key = "%d>%d" % (r["cin"], r["cout"])
    g[key] = g.get(key, 0.0) + r["ms"] * 1e3
print("conv us", round(sum(g.values())), " ".join("%s:%.0f" % kv for kv in g.items()))')"
done
