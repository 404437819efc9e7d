#!/bin/bash
# Final verification of a build on one B200 (run under gpurun): GPU tests, smoke, the default bench line, the reference arm,
# the other workloads, then the ncu passes (each after its plain command exited 0).
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; tail -3 gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -2 gpurun_out/final_smoke.log
( time python bench.py --kernels-out gpurun_out/final_kernels_nusc_LC.json > gpurun_out/final_bench_nusc_LC.json 2> gpurun_out/final_bench.err ) 2>&1 | tail -3
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/final_bench_reference_arm.json 2> gpurun_out/final_ref.err
for w in waymo_L kitti_L; do python bench.py --workload $w --steps 10 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/final_bench_$w.json 2> gpurun_out/final_$w.err; done
B="python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --frames-in-flight 1 --profiler-range"
$B > gpurun_out/plain_graph.log 2>&1 && timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/r02_launches_raw.csv $B > gpurun_out/ncu_launches.log 2>&1
ONLY=${ONLY:-'srf_img_roi|srf_bev_roi|srf_spconv_f32|srf_nchw_to_rows|srf_layernorm|srf_index_mark_strided|srf_conv3x3_rows'}
python tools/ncu_frame.py --only "$ONLY" > gpurun_out/plain_frame2.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on \
  --profile-from-start off -f -o gpurun_out/r02_frame2 python tools/ncu_frame.py --only "$ONLY" > gpurun_out/ncu_frame2.log 2>&1
ncu -i gpurun_out/r02_frame2.ncu-rep --page raw --csv > gpurun_out/r02_ncu_frame2_raw.csv
rm -f gpurun_out/r02_frame2.ncu-rep
ls -la gpurun_out | tail -20
