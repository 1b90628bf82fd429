# NOTE: the S2S_GCONV_SPEC_BIG switch (compile-time plans of the 16x32 / 4-pixel geometry) existed only in the build this script measured;
# the plans were rejected (batch-128 step 1538 vs 1460 us) and removed, see profiles/r2_summary.md and the comment in gconv.cuh.
# A/B of the compile-time plans of the filled-GPU regime (S2S_GCONV_SPEC_BIG) + parity of the deferred head finalize
set -x
python -m pytest tests/test_gpu_model.py tests/test_gpu_training_api.py -x -q > gpurun_out/r2m_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2m_gpu_tests.log
for big in 0 1; do
  S2S_GCONV_SPEC_BIG=$big python tools/profile_model.py --batch 128 --steps 30 > gpurun_out/r2m_profile_b128_big$big.log 2>&1; head -4 gpurun_out/r2m_profile_b128_big$big.log
  S2S_GCONV_SPEC_BIG=$big python bench.py --steps 20 --warmup 5 --large-batch 0 --extras 0 --concurrent-models 8 --no-cpu-baseline > gpurun_out/r2m_bench_big$big.json 2> gpurun_out/r2m_bench_big$big.err
  python -c "
import json;d=json.loads(open('gpurun_out/r2m_bench_big$big.json').read().strip().splitlines()[-1])
print('big$big', d['ms_per_step'], {k:v.get('samples_per_s') for k,v in d['inference_c5'].items() if isinstance(v,dict)}, d['trial_batching']['value'])"
done
python tools/profile_model.py --batch 16 --steps 100 > gpurun_out/r2m_profile_b16.log 2>&1; head -1 gpurun_out/r2m_profile_b16.log
