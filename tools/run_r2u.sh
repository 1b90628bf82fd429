set -x
python -m pytest tests/test_gpu_model.py tests/test_gpu_ops.py -x -q > gpurun_out/r2u_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2u_gpu_tests.log
python tools/profile_model.py --batch 16 --steps 300 --precision tf32 > gpurun_out/r2u_profile_b16_tf32.log 2>&1; head -1 gpurun_out/r2u_profile_b16_tf32.log | cut -c90-200
python tools/profile_model.py --batch 128 --steps 30 --precision tf32 > gpurun_out/r2u_profile_b128_tf32.log 2>&1; head -1 gpurun_out/r2u_profile_b128_tf32.log | cut -c90-200
python tools/profile_model.py --filters 3 --n-blocks 5 --ct 5 --cin 1 --batch 16 --steps 50 --precision tf32 > gpurun_out/r2u_profile_gridmax_tf32.log 2>&1; head -1 gpurun_out/r2u_profile_gridmax_tf32.log | cut -c90-200
python tools/profile_model.py --filters 3 --n-blocks 5 --ct 5 --cin 1 --batch 16 --steps 50 > gpurun_out/r2u_profile_gridmax_fp32.log 2>&1; head -1 gpurun_out/r2u_profile_gridmax_fp32.log | cut -c90-200
