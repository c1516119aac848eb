#!/bin/bash
# tools/ab_dbg.sh: FIR diagnostics modes of the in-tree library (quick workload, serial mode): ms per step (= 2 launches)
cd "$(dirname "$0")/.."
for d in 0 4 8 16 20 2 6; do
  GAIS_OVERLAP=0 GAIS_FIR_DBG=$d python bench.py --channels 65536 --frames 65536 --steps 3 --warmup 1 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('GAIS_FIR_DBG=$d  fir %.3f ms' % d['roofline']['chain']['fir_ms_per_step'])"
done
