#!/bin/bash
# A/B of library variants on the RoI / interaction stage: frames/s of both workloads + BEV / image sampler kernel times
for so in srfdet_b200/csrc/libsrfdet_b200.so srfdet_b200/csrc/libsrfdet_b200_v*.so; do
  [ -f "$so" ] || continue
  export SRFDET_B200_LIB=$PWD/$so
  a=$(timeout 200 python bench.py --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c 'import sys,json; print(json.loads(sys.stdin.read())["value"])')
  b=$(timeout 200 python bench.py --workload nusc_LC --steps 30 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | python -c 'import sys,json; print(json.loads(sys.stdin.read())["value"])')
  timeout 200 python tools/frame_timeline.py LC > /tmp/tl.txt 2>&1
  r=$(grep "bev_roi" /tmp/tl.txt | awk '{s+=$2; n++} END {printf "%.1f", s/n}')
  i=$(grep "img_roi" /tmp/tl.txt | awk '{s+=$2; n++} END {printf "%.1f", s/n}')
  echo "$(basename $so) nusc_L $a fps  nusc_LC $b fps  bev_roi $r us  img_roi $i us"
done
