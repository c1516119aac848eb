mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 300 > gpurun_out/s2_pytest_full.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s2_pytest_full.log
timeout 900 python bench.py > gpurun_out/r2_bench_fused_a.json 2> gpurun_out/r2_bench_fused_a.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r2_bench_fused_a.json; tail -3 gpurun_out/r2_bench_fused_a.err
