# fused vs two-kernel chain at mid-size batches (where should the automatic choice switch?)
for ch in 9472 16384 32768; do
  for f in 1 0; do
    GAIS_FUSED=$f python bench.py --channels $ch --frames 131072 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline --no-other-configs --no-gather-check --no-two-kernel 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); r=d['roofline']
print('channels $ch GAIS_FUSED=$f: step %.3f ms  (%s, chain %s)' % (d['ms_per_step'], r['kernel'], d['config']['chain']))"
  done
done
