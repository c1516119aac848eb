#!/bin/bash
# A/B on the GPU box: tools/ab.sh variant... ; per variant the quick workload (65536 ch x 65536 samples) with
# the FIR/tracker overlap on and off.  Prints ms per step and the per-kernel sums.
cd "$(dirname "$0")/.."
q() { python bench.py --channels 65536 --frames 65536 --steps 4 --warmup 2 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; c=r['chain']
print('  step %.3f ms | fir %.3f trk %.3f post %.3f | solo %s | ok %d' % (d['ms_per_step'], c['fir_ms_per_step'], c['track_ms_per_step'], c['post_ms_per_step'], r.get('solo'), d['counters_rank0']['ok']))"; }
for v in "$@"; do
  lib=gnuais_b200/lib/variants/$v.so; [ "$v" = default ] && lib=gnuais_b200/lib/libgaisb200.so
  echo "== $v overlap"; GAIS_B200_LIB=$PWD/$lib q
  echo "== $v serial";  GAIS_B200_LIB=$PWD/$lib GAIS_OVERLAP=0 q
done
