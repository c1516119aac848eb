# tools/_run6.sh variant...: quick workload (65536 ch x 65536 samples) per library variant (default = the built library)
mkdir -p gpurun_out
q() { python bench.py --channels 65536 --frames 65536 --steps 4 --warmup 2 --no-e2e --no-cpu-baseline --no-other-configs --no-gather-check --no-two-kernel 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('  step %.3f ms | %s %.3f ms/launch | frac %.3f | ok %d' % (d['ms_per_step'], r['kernel'], r['avg_launch_ms'], r['frac'], d['counters_rank0']['ok']))"; }
for v in "$@"; do
  lib=$PWD/gnuais_b200/lib/variants/$v.so; [ "$v" = default ] && lib=$PWD/gnuais_b200/lib/libgaisb200.so
  echo "== $v"; GAIS_B200_LIB=$lib q
done
