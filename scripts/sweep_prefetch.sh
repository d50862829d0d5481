#!/bin/bash
# GPU box: parity first, then the draft+verify step time for several pre-wait L2 prefetch budgets.
set -e
timeout 300 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for mb in -1 8 16 32 64 128; do
  DFLASH_PREFETCH_MB=$mb python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>&1 | tail -1 |
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('prefetch_mb $mb', d['step_us'], round(d['step_roofline']['frac'],4))"
done
