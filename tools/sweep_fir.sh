run() { echo "== $*"; env "$@" timeout 120 python bench.py --channels 65536 --frames 65536 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['roofline']['chain'], d['counters_rank0'])"; }
run GAIS_FIR_DBG=0
run GAIS_FIR_DBG=3
for t in 8192 16384 32768 65536; do run GAIS_TILE_FRAMES=$t; run GAIS_TILE_FRAMES=$t GAIS_FIR_SPB=16; done
