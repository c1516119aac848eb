set -x
mkdir -p gpurun_out
timeout 180 python __graft_entry__.py smoke > gpurun_out/s2_smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/s2_smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 240 > gpurun_out/s2_pytest2.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/s2_pytest2.log
timeout 300 bash tools/ab.sh default > gpurun_out/s2_ab_fused.txt 2>&1; cat gpurun_out/s2_ab_fused.txt
