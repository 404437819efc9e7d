#!/bin/bash
# Round profile on one B200 (run under gpurun): every ncu pass follows a plain run of the same command that exited 0.
#   1. launch list of the timed region of bench.py (graph replay, one frame in flight), gpu__time_duration only
#   2. ncu --set full of ONE eagerly launched frame, restricted to the first call of every distinct C-ABI signature
# Raw csv goes to gpurun_out/ (scratch); tools/ncu_summary.py turns it into the summaries committed under profiles/.
set -x
mkdir -p gpurun_out
B="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --frames-in-flight 1 --profiler-range"
$B > gpurun_out/plain_graph.log 2>&1 || { tail -5 gpurun_out/plain_graph.log; exit 1; }
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r02_launches_raw.csv \
  $B > gpurun_out/ncu_launches.log 2>&1
python tools/ncu_frame.py > gpurun_out/plain_frame.log 2>&1 || { tail -5 gpurun_out/plain_frame.log; exit 1; }
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o gpurun_out/r02_frame \
  python tools/ncu_frame.py > gpurun_out/ncu_frame.log 2>&1
ncu -i gpurun_out/r02_frame.ncu-rep --page raw --csv > gpurun_out/r02_ncu_frame_raw.csv
ls -la gpurun_out/r02_frame.ncu-rep
if [ $(stat -c %s gpurun_out/r02_frame.ncu-rep) -gt 30000000 ]; then rm -f gpurun_out/r02_frame.ncu-rep; fi
tail -3 gpurun_out/plain_graph.log gpurun_out/ncu_frame.log
ls -la gpurun_out | tail -12
