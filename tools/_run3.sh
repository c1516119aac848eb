set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"ais_fused_kernel" --launch-skip 1 --launch-count 1 -o gpurun_out/s2_fused1 -f python bench.py --channels 65536 --frames 32768 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-other-configs --no-gather-check > gpurun_out/s2_ncu3.log 2>&1; tail -3 gpurun_out/s2_ncu3.log | cut -c1-300
