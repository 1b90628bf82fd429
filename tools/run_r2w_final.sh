# final validation of the build: the whole GPU suite, both bench arms as the driver runs them, then the evidence captures
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2w_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2w_gpu_tests.log
python bench.py --impl reference > gpurun_out/r2w_bench_reference_arm.json 2> gpurun_out/r2w_bench_reference_arm.err; echo "ref rc=$?"
python bench.py > gpurun_out/r2w_bench_1gpu.json 2> gpurun_out/r2w_bench_1gpu.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/r2w_bench_1gpu.json
CMD="python bench.py --steps 4 --warmup 3 --large-batch 0 --inference-c5 0 --extras 0 --concurrent-models 0 --no-cpu-baseline --profile-steps 1"
$CMD > gpurun_out/r2w_plain.log 2> gpurun_out/r2w_plain.err || { tail -5 gpurun_out/r2w_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r2w_launches_b16.csv $CMD > gpurun_out/r2w_ncu_list.log 2>&1
tail -2 gpurun_out/r2w_ncu_list.log | cut -c1-200
ncu --set full --import-source on --clock-control none -k regex:"gconv_kernel" --launch-skip 120 -c 12 -o /tmp/r2w_gconv $CMD > gpurun_out/r2w_ncu_full.log 2>&1
ncu -i /tmp/r2w_gconv.ncu-rep --page raw --csv > gpurun_out/r2w_ncu_full_gconv_b16_raw.csv 2>/dev/null
ls -la gpurun_out/r2w_*
