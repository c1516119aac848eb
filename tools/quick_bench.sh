# quick A/B: 65536 channels x 65536 samples (4.29 Gsamples per step); prints per-stage ms
env "$@" timeout 120 python bench.py --channels 65536 --frames 65536 --steps 3 --warmup 2 --no-e2e --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['roofline']['chain'], d['counters_rank0'])"
