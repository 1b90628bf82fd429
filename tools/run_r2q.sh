set -x
python -m pytest tests/test_gpu_model.py tests/test_gpu_training_api.py -x -q > gpurun_out/r2q_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2q_gpu_tests.log
python tools/profile_model.py --batch 16 --steps 300 > gpurun_out/r2q_profile_b16.log 2>&1; head -1 gpurun_out/r2q_profile_b16.log
python tools/profile_model.py --batch 128 --steps 30 > gpurun_out/r2q_profile_b128.log 2>&1; head -1 gpurun_out/r2q_profile_b128.log
python tools/profile_model.py --batch 16 --steps 300 --precision tf32 > gpurun_out/r2q_profile_b16_tf32.log 2>&1; head -1 gpurun_out/r2q_profile_b16_tf32.log
