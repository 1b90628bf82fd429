set -x
CMD="python bench.py --steps 4 --warmup 3 --large-batch 0 --inference-c5 0 --concurrent-models 0 --no-cpu-baseline --profile-steps 1"
$CMD > gpurun_out/plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r1b_launches.csv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'wgrad_kernel|gconv_kernel' -s 60 -c 24 -o gpurun_out/r1b_prof $CMD > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out
