# round-2 (session 3) diagnostics at batch 128, fp32: per-tag table, launch list, full captures of wgrad / gconv of one step
set -x
CMD="python tools/profile_model.py --batch 128 --steps 2"
$CMD > gpurun_out/r2k_plain.log 2> gpurun_out/r2k_plain.err || { tail -5 gpurun_out/r2k_plain.err; exit 1; }
cat gpurun_out/r2k_plain.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2k_launches_b128.csv $CMD > gpurun_out/r2k_ncu_list.log 2>&1
tail -2 gpurun_out/r2k_ncu_list.log | cut -c1-200
# one step = 74 launches; 5 warm-up + 2 timed graph replays + eager: skip the first 7 steps' worth of each family
ncu --set full --import-source on --clock-control none -k regex:"wgrad_kernel" --launch-skip 119 -c 17 -o /tmp/r2k_wgrad $CMD > gpurun_out/r2k_ncu_full_wgrad.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"gconv_kernel" --launch-skip 210 -c 30 -o /tmp/r2k_gconv $CMD > gpurun_out/r2k_ncu_full_gconv.log 2>&1
ls -la /tmp/r2k_*.ncu-rep
for f in wgrad gconv; do
  ncu -i /tmp/r2k_$f.ncu-rep --page raw --csv > gpurun_out/r2k_ncu_full_${f}_b128_raw.csv 2>/dev/null
  sz=$(stat -c %s /tmp/r2k_$f.ncu-rep)
  if [ "$sz" -lt 28000000 ]; then cp /tmp/r2k_$f.ncu-rep gpurun_out/; fi
done
ls -la gpurun_out/r2k_*
