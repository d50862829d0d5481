timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2t_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2t_tests.log
B="--no-cpu-baseline --no-sharded --no-gpu-reference --no-full-cycle"
python bench.py --steps 200 --warmup 20 $B > gpurun_out/r2t_b1.json 2>gpurun_out/r2t_b1.err
for R in 8 16 32 64; do
python bench.py --requests $R --steps 40 --warmup 5 $B > gpurun_out/r2t_wide_b$R.json 2>gpurun_out/r2t_wide_b$R.err
DFLASH_FUSED_MIN_ROWS=100000 python bench.py --requests $R --steps 40 --warmup 5 $B > gpurun_out/r2t_part_b$R.json 2>gpurun_out/r2t_part_b$R.err
done
DFLASH_FUSED_MIN_ROWS=64 python bench.py --requests 8 --steps 40 --warmup 5 $B > gpurun_out/r2t_wide64_b8.json 2>/dev/null
DFLASH_FUSED_MIN_ROWS=64 python bench.py --requests 4 --steps 40 --warmup 5 $B > gpurun_out/r2t_wide64_b4.json 2>/dev/null
python bench.py --requests 4 --steps 40 --warmup 5 $B > gpurun_out/r2t_part_b4.json 2>/dev/null
python -c "
import json
for v in ('b1','part_b4','wide64_b4','wide_b8','part_b8','wide64_b8','wide_b16','part_b16','wide_b32','part_b32','wide_b64','part_b64'):
    try:
        d=json.load(open('gpurun_out/r2t_%s.json'%v)); print(v, d['step_us'], round(d['value']), round(d['e2e']['value']), d['launches_per_step'])
    except Exception as e: print(v,'ERR',e)"
