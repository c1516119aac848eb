mkdir -p gpurun_out
A="--no-e2e --no-cpu-baseline --no-other-configs --no-gather-check --no-two-kernel"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 $A > gpurun_out/r2_launches_bench.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"ais_fused_kernel" --launch-skip 1 --launch-count 1 -o gpurun_out/r2_fused_final -f python bench.py --steps 1 --warmup 1 $A > gpurun_out/r2_ncu_fused_final.log 2>&1; echo "ncu rc=$?"
