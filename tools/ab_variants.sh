#!/bin/bash
# A/B runs of development variants of the library (python -m srfdet_b200.build --variant NAME MACRO=VALUE ...)
# usage: tools/ab_variants.sh [bench args] ; prints frames/s and the sparse-conv time per channel group per variant
for so in srfdet_b200/csrc/libsrfdet_b200.so srfdet_b200/csrc/libsrfdet_b200_v*.so; do
  [ -f "$so" ] || continue
  out=$(SRFDET_B200_LIB=$PWD/$so timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --layers-out /tmp/layers.json "$@" 2>&1 | tail -1)
  echo "$(basename $so) $(echo "$out" | python -c '
import sys, json, collections
d = json.loads(sys.stdin.read())
rows = json.load(open("/tmp/layers.json"))
rows = rows["layers"] if isinstance(rows, dict) else rows
g = collections.OrderedDict()
for r in rows:
    key = "%d>%d" % (r["cin"], r["cout"])
    g[key] = g.get(key, 0.0) + r["ms"] * 1e3
print(d["value"], "fps  conv us", round(sum(g.values())), " ".join("%s:%.0f" % kv for kv in g.items()))' 2>&1 | tail -1)"
done
