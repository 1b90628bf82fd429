# round-2 launch list + full capture of the step's dominant kernel (run under gpurun, one GPU).
#   1. plain run of the short bench (must exit 0), 2. gpu__time_duration launch list of the same command,
#   3. ncu --set full of the first gconv launches of a train step (raw + source CSV pages; the .ncu-rep stays on the box)
set -x
CMD="python bench.py --steps 4 --warmup 3 --large-batch 0 --inference-c5 0 --extras 0 --concurrent-models 0 --no-cpu-baseline --profile-steps 1"
$CMD > gpurun_out/r2f_plain.log 2> gpurun_out/r2f_plain.err || { tail -5 gpurun_out/r2f_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2f_launches_b16.csv $CMD > gpurun_out/r2f_ncu_list.log 2>&1
tail -2 gpurun_out/r2f_ncu_list.log | cut -c1-200
ncu --set full --import-source on --clock-control none -k regex:"gconv_kernel" --launch-skip 120 -c 12 -o /tmp/r2f_gconv $CMD > gpurun_out/r2f_ncu_full.log 2>&1
ncu -i /tmp/r2f_gconv.ncu-rep --page raw --csv > gpurun_out/r2f_ncu_full_gconv_raw.csv 2>/dev/null
ls -la gpurun_out/r2f_*
