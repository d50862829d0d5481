#!/bin/bash
# Timing ablation of the draft+verify step (results are wrong when kernels are skipped; timing only).
# usage: scripts/ablate.sh [requests] [masks...]     mask bits: 1 finalize_rows, 2 swiglu, 4 attn_combine,
#        8 qkv_post, 16 attn_split, 32 partial GEMMs, 64 lm_head GEMM
R=${1:-1}; shift
MASKS=${@:-0 1 2 4 8 16 31 32 64 96 127}
for m in $MASKS; do
  DFLASH_DEBUG_SKIP=$m python bench.py --steps 60 --warmup 6 --requests $R --no-cpu-baseline --no-full-cycle 2>&1 | tail -1 |
    python -c "import json,sys; d=json.loads(sys.stdin.read()); print('R=$R skip mask $m', d['step_us'])" 2>&1 | tail -1
done
