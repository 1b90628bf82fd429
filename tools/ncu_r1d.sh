# final launch list of round 1 (r1d): every launch of a short default bench run with its device time
set -x
CMD="python bench.py --steps 4 --warmup 3 --large-batch 0 --inference-c5 0 --concurrent-models 0 --no-cpu-baseline --profile-steps 1"
$CMD > gpurun_out/plain_d.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r1d_launches.csv $CMD > gpurun_out/ncu_d.log 2>&1
ncu --set full --clock-control none -k regex:'wgrad_kernel|acc_kernel|bn_apply_kernel' -s 40 -c 12 -o /tmp/r1d_prof $CMD > gpurun_out/ncu_d2.log 2>&1
ncu -i /tmp/r1d_prof.ncu-rep --page raw --csv > gpurun_out/r1d_raw.csv 2>/dev/null
ls -la gpurun_out | tail -6
