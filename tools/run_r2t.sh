set -x
python -m pytest tests/test_gpu_model.py tests/test_gpu_training_api.py tests/test_gpu_ops.py -x -q > gpurun_out/r2t_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2t_gpu_tests.log
for rep in 1 2; do for ne in 1 0; do
  S2S_NO_EARLY_LOADS=$ne python tools/profile_model.py --batch 16 --steps 400 > gpurun_out/r2t_profile_b16_noearly${ne}_rep$rep.log 2>&1; echo "noearly=$ne $(head -1 gpurun_out/r2t_profile_b16_noearly${ne}_rep$rep.log | cut -c100-180)"
done; done
for ne in 1 0; do
S2S_NO_EARLY_LOADS=$ne python bench.py --steps 50 --warmup 5 --large-batch 0 --extras 0 --inference-c5 0 --concurrent-models 8 --no-cpu-baseline > gpurun_out/r2t_bench_noearly$ne.json 2>/dev/null
python -c "
import json
d=json.loads(open('gpurun_out/r2t_bench_noearly$ne.json').read().strip().splitlines()[-1]); print('noearly=$ne bench', d['ms_per_step'], d['trial_batching']['value'])"
done
