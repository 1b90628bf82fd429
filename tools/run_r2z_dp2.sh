# 2 GPUs, end-of-round build: data-parallel parity tests and the bench line
set -x
python -m pytest tests/test_gpu_dp_peer.py tests/test_gpu_dp.py -x -q > gpurun_out/r2z_dp2_tests.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2z_dp2_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 > gpurun_out/r2z_bench_dp2.json 2> gpurun_out/r2z_bench_dp2.err; echo "bench rc=$?"
cut -c1-300 gpurun_out/r2z_bench_dp2.json
