mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_fused_gpu.py tests/test_parity_gpu.py -m gpu -x -q --timeout 400 > gpurun_out/s2_pytest15.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/s2_pytest15.log
bash tools/_run6.sh default lay0
