mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -x -q --timeout 240 > gpurun_out/s2_pytest7.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/s2_pytest5.log
q() { python bench.py --channels 65536 --frames 65536 --steps 4 --warmup 2 --no-e2e --no-cpu-baseline --no-other-configs --no-gather-check 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; c=r['chain']
print('  step %.3f ms | fir %.3f trk %.3f post %.3f | ok %d' % (d['ms_per_step'], c['fir_ms_per_step'], c['track_ms_per_step'], c['post_ms_per_step'], d['counters_rank0']['ok']))"; }
q
timeout 900 ncu --set full --clock-control none --import-source on --kernel-name regex:"ais_fused_kernel" --launch-skip 1 --launch-count 1 -o gpurun_out/s2_fused5 -f python bench.py --channels 65536 --frames 32768 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-other-configs --no-gather-check > gpurun_out/s2_ncu7.log 2>&1
