# validation of a build: GPU parity tests, per-tag profiles at batch 16 / 128, the default bench line
set -x
python -m pytest tests -m gpu -x -q > gpurun_out/r2l_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2l_gpu_tests.log
python tools/profile_model.py --batch 16 --steps 50 > gpurun_out/r2l_profile_b16.log 2>&1; head -1 gpurun_out/r2l_profile_b16.log
python tools/profile_model.py --batch 128 --steps 20 > gpurun_out/r2l_profile_b128.log 2>&1; head -1 gpurun_out/r2l_profile_b128.log
python bench.py > gpurun_out/r2l_bench.json 2> gpurun_out/r2l_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2l_bench.json
