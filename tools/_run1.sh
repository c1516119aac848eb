set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s2_pytest.log 2>&1; tail -3 gpurun_out/s2_pytest.log
bash tools/ab.sh default > gpurun_out/s2_ab_default.txt 2>&1; cat gpurun_out/s2_ab_default.txt
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name regex:"fir_sign_tc_kernel|track_kernel" --launch-skip 2 --launch-count 2 -o gpurun_out/s2_cur -f python bench.py --channels 65536 --frames 49152 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-other-configs --no-gather-check > gpurun_out/s2_ncu.log 2>&1; tail -3 gpurun_out/s2_ncu.log
