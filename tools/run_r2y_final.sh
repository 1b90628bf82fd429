# end-of-round validation: the whole GPU suite and the default bench line (as the driver runs them), tf32 spot checks
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2y_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2y_gpu_tests.log
python bench.py > gpurun_out/r2y_bench_1gpu.json 2> gpurun_out/r2y_bench_1gpu.err; echo "bench rc=$?"
python -c "
import json
d=json.loads(open('gpurun_out/r2y_bench_1gpu.json').read().strip().splitlines()[-1]); print('bench', d['value'], d['ms_per_step'], d['e2e']['value'], {k:v.get('samples_per_s') for k,v in d['inference_c5'].items() if isinstance(v,dict)}, d['large_batch']['value'], d['large_batch']['tf32']['samples_per_s'], d['trial_batching']['value'], d['roofline']['frac'])"
python tools/profile_model.py --batch 16 --steps 300 | head -1 | cut -c90-200
python tools/profile_model.py --batch 16 --steps 300 --precision tf32 | head -1 | cut -c90-200
python -c "import __graft_entry__ as g; g.smoke()"
