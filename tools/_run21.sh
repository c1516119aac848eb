mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_fused_gpu.py tests/test_parity_gpu.py -m gpu -x -q --timeout 400 > gpurun_out/s2_pytest21.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/s2_pytest21.log
bash tools/_run6.sh default
timeout 800 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_fused_gpu.py -m gpu -x -q --timeout 700 -k "33-1000" > gpurun_out/r2_racecheck.log 2>&1; echo "racecheck rc=$?"; grep -E "RACECHECK SUMMARY|Race reported|passed|failed" gpurun_out/r2_racecheck.log | head
