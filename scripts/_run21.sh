B="--no-cpu-baseline --no-sharded --no-gpu-reference --no-full-cycle"
for sp in 8 12 16; do
DFLASH_ATTN_SPLITS=$sp python bench.py --steps 200 --warmup 20 $B > gpurun_out/r3a_sp$sp.json 2>/dev/null
done
for R in 4 8; do
python bench.py --requests $R --steps 100 --warmup 10 $B > gpurun_out/r3a_rb32_b$R.json 2>/dev/null
DFLASH_LIB=$PWD/build/libdflash_rb128.so python bench.py --requests $R --steps 100 --warmup 10 $B > gpurun_out/r3a_rb128_b$R.json 2>/dev/null
done
python -c "
import json
for v in ('sp8','sp12','sp16','rb32_b4','rb128_b4','rb32_b8','rb128_b8'):
    try:
        d=json.load(open('gpurun_out/r3a_%s.json'%v)); print(v, d['step_us'], round(d['value']), round(d['e2e']['value']), d['launches_per_step'])
    except Exception as e: print(v,'ERR',e)"
