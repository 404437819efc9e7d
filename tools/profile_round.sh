set -x
B="python bench.py --steps 1 --warmup 3 --no-extras --no-cpu-baseline --no-graph --frames-in-flight 1"
$B > gpurun_out/plain_nograph.log 2>&1 || exit 1
python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --frames-in-flight 1 > gpurun_out/plain_graph.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_launches_raw.csv python bench.py --steps 2 --warmup 3 --no-extras --no-cpu-baseline --frames-in-flight 1 > gpurun_out/ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'igemm_umma_kernel|conv3x3_halo|spconv16_warp|spconv_smallcin' --launch-skip 230 -c 75 -o gpurun_out/r02_ncu_conv $B > gpurun_out/ncu_conv.log 2>&1
ncu -i gpurun_out/r02_ncu_conv.ncu-rep --page raw --csv > gpurun_out/r02_ncu_conv_raw.csv
ncu --set full --clock-control none --import-source on -k regex:'img_roi|bev_roi|mha_attention|dynconv_interact|dwconv3x3|rulebook|index_emit|index_mark|hv_|layernorm|linear_smalln|gemv|channel_sum|scan_' --launch-skip 300 -c 110 -o gpurun_out/r02_ncu_others $B > gpurun_out/ncu_others.log 2>&1
ncu -i gpurun_out/r02_ncu_others.ncu-rep --page raw --csv > gpurun_out/r02_ncu_others_raw.csv
rm -f gpurun_out/r02_ncu_conv.ncu-rep gpurun_out/r02_ncu_others.ncu-rep
ls -la gpurun_out | tail -12
