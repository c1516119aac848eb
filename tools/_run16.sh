mkdir -p gpurun_out
A="--no-e2e --no-cpu-baseline --no-other-configs --no-gather-check --no-two-kernel"
timeout 900 python bench.py --steps 5 --warmup 3 $A 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('FULL step %.3f ms | %s %.3f ms/launch | frac %.3f | ok %d | clocks %s' % (d['ms_per_step'], r['kernel'], r['avg_launch_ms'], r['frac'], d['counters_rank0']['ok'], d['clocks']))"
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"ais_fused_kernel" --launch-skip 1 --launch-count 1 -o gpurun_out/s2_fused_nf -f python bench.py --channels 65536 --frames 65536 --steps 1 --warmup 1 $A > gpurun_out/s2_ncu_nf.log 2>&1; echo "ncu rc=$?"
