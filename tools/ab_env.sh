#!/bin/bash
# tools/ab_env.sh variant "ENV=1 ENV2=2" ... : quick workload, serial mode (kernels timed alone)
cd "$(dirname "$0")/.."
v=$1; shift
lib=$PWD/gnuais_b200/lib/variants/$v.so
for e in "$@"; do
  echo "== $v $e"
  env GAIS_B200_LIB=$lib GAIS_OVERLAP=0 $e python bench.py --channels 65536 --frames 65536 --steps 3 --warmup 1 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); c=d['roofline']['chain']
print('  step %.3f ms | fir %.3f trk %.3f | ok %d' % (d['ms_per_step'], c['fir_ms_per_step'], c['track_ms_per_step'], d['counters_rank0']['ok']))"
done
