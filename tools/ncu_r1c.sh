# ncu --set full of the conv-family kernels at the throughput size (batch 128), r1c
set -x
CMD="python bench.py --batch 128 --steps 3 --warmup 3 --large-batch 0 --inference-c5 0 --concurrent-models 0 --no-cpu-baseline --profile-steps 1"
$CMD > gpurun_out/plain_c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'wgrad_kernel|gconv_kernel' -s 141 -c 47 -o gpurun_out/r1c_prof $CMD > gpurun_out/ncu_c.log 2>&1
ls -la gpurun_out | tail -5
