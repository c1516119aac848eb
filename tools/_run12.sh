mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_dist_peer_gpu.py -m gpu -x -q --timeout 300 > gpurun_out/r2_pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_2gpu_fused.json 2> gpurun_out/r2_bench_2gpu_fused.err; echo "bench rc=$?"; tail -c 1500 gpurun_out/r2_bench_2gpu_fused.json | head -c 1500; tail -2 gpurun_out/r2_bench_2gpu_fused.err
