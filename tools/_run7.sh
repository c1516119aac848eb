mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q --timeout 240 > gpurun_out/s2_pytest9.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s2_pytest9.log
bash tools/_run6.sh "$@"
