#!/bin/bash
# tools/ab_last.sh "ENV=…" … : the default workload under different environments (overlap on): ms per step and solo launch times
cd "$(dirname "$0")/.."
for e in "$@"; do env $e python bench.py --steps 5 --warmup 2 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']; print('$e', 'step %.3f ms' % d['ms_per_step'], 'solo trk %.4f fir %.4f' % (r['solo']['track_launch_ms'], r['solo']['fir_launch_ms']))"; done
