#!/bin/bash
# time-tile length sweep on the default workload: tools/sweep_tiles.sh [tile ...]
cd "$(dirname "$0")/.."
for t in ${@:-16384 24576 32768 49152 65536}; do
  python bench.py --steps 5 --warmup 2 --no-e2e --no-cpu-baseline --tile-frames $t 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('tile %6d  step %.3f ms  %s' % (d['config']['tile_frames'], d['ms_per_step'], d['roofline']['tile_plan_frames']))"
done
