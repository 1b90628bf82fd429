# ncu --set full of the conv-family kernels at the throughput size (batch 128), r1c; the report is reduced to CSV on the box
set -x
CMD="python bench.py --batch 128 --steps 3 --warmup 3 --large-batch 0 --inference-c5 0 --concurrent-models 0 --no-cpu-baseline --profile-steps 1"
$CMD > gpurun_out/plain_c.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'wgrad_kernel|gconv_kernel' -s 141 -c 47 -o /tmp/r1c_prof $CMD > gpurun_out/ncu_c.log 2>&1
ncu -i /tmp/r1c_prof.ncu-rep --page raw --csv > gpurun_out/r1c_raw.csv 2>/dev/null
for i in 0 1 2 3 4 5 6 7 8 9 10 11 12 13; do
  ncu -i /tmp/r1c_prof.ncu-rep --page source --csv --print-source cuda,sass --launch-skip $i --launch-count 1 2>/dev/null | gzip > gpurun_out/r1c_src_$i.csv.gz
done
ls -la gpurun_out | tail -20
