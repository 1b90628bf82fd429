# round-2 ncu capture (run under gpurun, one GPU): --set full of the convolution kernels, exported as raw CSV.
# The .ncu-rep itself (>100 MB with source import) is deleted on the box: gpurun_out/ is limited to 64 MiB.
set -x
CMD="python tools/ncu_targets.py 1"
$CMD > gpurun_out/ncu_r2_plain.log 2>&1 && \
ncu --set full --clock-control none -k regex:"tc3conv_kernel|tcconv_kernel|gconv_kernel" -c 50 -o /tmp/r2_conv_full $CMD > gpurun_out/ncu_r2_full.log 2>&1
tail -3 gpurun_out/ncu_r2_full.log
ncu -i /tmp/r2_conv_full.ncu-rep --page raw --csv > gpurun_out/r2_conv_full_raw.csv 2>/dev/null
ls -la /tmp/r2_conv_full.ncu-rep gpurun_out/r2_conv_full_raw.csv
