# launch list of the final round-1 build (r1e)
CMD="python bench.py --steps 4 --warmup 3 --large-batch 0 --inference-c5 0 --concurrent-models 0 --no-cpu-baseline --profile-steps 1"
$CMD > gpurun_out/plain_e.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1e_launches.csv $CMD > gpurun_out/ncu_e.log 2>&1
tail -2 gpurun_out/ncu_e.log | cut -c1-200
