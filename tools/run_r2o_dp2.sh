# 2 GPUs: data-parallel parity tests, the bench line, and the NVLink byte counters around a fixed number of steps
set -x
python -m pytest tests/test_gpu_dp_peer.py tests/test_gpu_dp.py -x -q > gpurun_out/r2o_dp2_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2o_dp2_tests.log
nvidia-smi nvlink -gt d > gpurun_out/r2o_nvlink_before.txt 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 2000 --warmup 10 --no-cpu-baseline --large-batch 0 --inference-c5 0 --extras 0 --concurrent-models 0 > gpurun_out/r2o_bench_dp2_2000steps.json 2> gpurun_out/r2o_bench_dp2_2000steps.err; echo "bench rc=$?"
nvidia-smi nvlink -gt d > gpurun_out/r2o_nvlink_after.txt 2>&1
cut -c1-600 gpurun_out/r2o_bench_dp2_2000steps.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > gpurun_out/r2o_bench_dp2.json 2> gpurun_out/r2o_bench_dp2.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/r2o_bench_dp2.json
head -30 gpurun_out/r2o_nvlink_after.txt
