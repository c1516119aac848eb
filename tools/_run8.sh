mkdir -p gpurun_out
GAIS_B200_LIB=$PWD/gnuais_b200/lib/variants/$1.so timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"ais_fused_kernel" --launch-skip 1 --launch-count 1 -o gpurun_out/s2_$1 -f python bench.py --channels 65536 --frames 32768 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-other-configs --no-gather-check > gpurun_out/s2_ncu_$1.log 2>&1
