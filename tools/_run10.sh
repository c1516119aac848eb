mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_fused_gpu.py tests/test_parity_gpu.py -m gpu -x -q --timeout 300 > gpurun_out/s2_pytest_fused.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/s2_pytest_fused.log
