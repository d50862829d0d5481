#!/bin/bash
# Timing ablation of the draft+verify step (results are wrong when kernels are skipped; timing only).
for m in 0 1 2 4 8 16 31 32 64 96 127; do
  DFLASH_DEBUG_SKIP=$m python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-full-cycle 2>&1 | tail -1 |
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('skip mask $m', d['step_us'])" 2>&1 | tail -1
done
