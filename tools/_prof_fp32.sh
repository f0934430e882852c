set -x
python tools/bench_layer.py m2g prec=fp32 > gpurun_out/fp32_m2g_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"rowmlp_tc_(dgrad|wgrad|fwd)_kernel" --launch-skip 38 -c 6 -o gpurun_out/prof_r2_fp32_m2g -f python tools/bench_layer.py m2g prec=fp32 > gpurun_out/fp32_m2g_ncu.log 2>&1
tail -3 gpurun_out/fp32_m2g_ncu.log
