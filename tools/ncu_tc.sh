# ncu capture (run under gpurun, one GPU) of the tcgen05 kernels on a thick layer (b16 64x64 96->96) and a thin one:
# tensor-pipe utilisation, shared-memory traffic and issue stalls.  Raw CSV pages only (the .ncu-rep stays on the box).
set -x
tools/bin/tcwg_test time 1 0 1 > gpurun_out/ncu_tc_plain.log 2>&1 || exit 1
ncu --set full --import-source on --clock-control none -k regex:"tcwgrad_kernel" --launch-skip 1030 -c 2 -o /tmp/tcwg_full tools/bin/tcwg_test time 1 0 1 > gpurun_out/ncu_tc_a.log 2>&1
ncu -i /tmp/tcwg_full.ncu-rep --page raw --csv > gpurun_out/r2d_ncu_tcwgrad_thick_raw.csv 2>/dev/null
ncu -i /tmp/tcwg_full.ncu-rep --page source --csv > gpurun_out/r2d_ncu_tcwgrad_thick_source.csv 2>/dev/null
ncu --set full --clock-control none -k regex:"tcwgrad_kernel" --launch-skip 300 -c 2 -o /tmp/tcwg_thin tools/bin/tcwg_test time 1 0 1 > gpurun_out/ncu_tc_b.log 2>&1
ncu -i /tmp/tcwg_thin.ncu-rep --page raw --csv > gpurun_out/r2d_ncu_tcwgrad_thin_raw.csv 2>/dev/null
ncu --set full --import-source on --clock-control none -k regex:"tc3conv_kernel" --launch-skip 500 -c 2 -o /tmp/tc3_full tools/bin/tc3_test 0 time 20 > gpurun_out/ncu_tc_c.log 2>&1
ncu -i /tmp/tc3_full.ncu-rep --page raw --csv > gpurun_out/r2d_ncu_tc3conv_thick_raw.csv 2>/dev/null
ncu -i /tmp/tc3_full.ncu-rep --page source --csv > gpurun_out/r2d_ncu_tc3conv_thick_source.csv 2>/dev/null
ls -la gpurun_out/r2d_*
