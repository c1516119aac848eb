#!/bin/bash
# tools/ab_full.sh variant... : the default workload (65536 ch x 480000 samples), overlap on; ms per step and solo launch times
cd "$(dirname "$0")/.."
for v in "$@"; do
  lib=$PWD/gnuais_b200/lib/variants/$v.so; [ "$v" = default ] && lib=$PWD/gnuais_b200/lib/libgaisb200.so
  echo "== $v"
  GAIS_B200_LIB=$lib python bench.py --steps 4 --warmup 2 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; c=r['chain']; s=r.get('solo') or {}
print('  step %.3f ms | in-step fir %.3f trk %.3f | solo fir %.4f trk %.4f ms/launch (fir frac %.3f) | ok %d' % (d['ms_per_step'], c['fir_ms_per_step'], c['track_ms_per_step'], s.get('fir_launch_ms',0), s.get('track_launch_ms',0), s.get('fir_frac',0), d['counters_rank0']['ok']))"
done
