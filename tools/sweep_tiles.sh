cd /root/repo
for t in 16384 24576 32768 49152 65536; do
  echo "== tile $t"; GAIS_TILE_FRAMES=$t python bench.py --steps 4 --warmup 2 --no-e2e --no-cpu-baseline --tile-frames $t 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('  step %.3f ms' % d['ms_per_step'], d['config']['tile_frames'], d['counters_rank0']['ok'])"
done
