# cfg2 / cfg3 (the latency-bound small configurations) per library variant
for v in "$@"; do
  lib=$PWD/gnuais_b200/lib/variants/$v.so; [ "$v" = default ] && lib=$PWD/gnuais_b200/lib/libgaisb200.so
  echo "== $v"; GAIS_B200_LIB=$lib python bench.py --channels 4096 --frames 65536 --steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-gather-check --no-two-kernel 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
for k,v in d['other_configs'].items(): print('  %s: %.3f ms (fir %.3f, track %.3f)' % (k[:5], v['ms_per_run'], v['fir_ms'], v['track_ms']))"
done
