# parity + A/B of the BatchNorm-backward statistics fold (S2S_NO_BN_FOLD=1 = bn_bwd_reduce kernels as before)
set -x
python -m pytest tests/test_gpu_model.py tests/test_gpu_training_api.py tests/test_gpu_ops.py -x -q > gpurun_out/r2n_gpu_tests.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2n_gpu_tests.log
for nf in 1 0; do  # 0 = fold on
  S2S_NO_BN_FOLD=$nf python tools/profile_model.py --batch 16 --steps 200 > gpurun_out/r2n_profile_b16_nofold$nf.log 2>&1; head -1 gpurun_out/r2n_profile_b16_nofold$nf.log
  S2S_NO_BN_FOLD=$nf python tools/profile_model.py --batch 128 --steps 30 > gpurun_out/r2n_profile_b128_nofold$nf.log 2>&1; head -1 gpurun_out/r2n_profile_b128_nofold$nf.log
done
S2S_NO_BN_FOLD_POOL=1 python tools/profile_model.py --batch 16 --steps 200 > gpurun_out/r2n_profile_b16_nopool.log 2>&1; head -1 gpurun_out/r2n_profile_b16_nopool.log
