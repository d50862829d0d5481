B="--no-cpu-baseline --no-sharded --no-gpu-reference --no-full-cycle"
for R in 16 32 64; do
python bench.py --requests $R --steps 40 --warmup 5 $B > gpurun_out/r3e_base_b$R.json 2>/dev/null
DFLASH_LIB=$PWD/build/libdflash_rbALL.so python bench.py --requests $R --steps 40 --warmup 5 $B > gpurun_out/r3e_rbALL_b$R.json 2>/dev/null
done
python -c "
import json
for v in ('base_b16','rbALL_b16','base_b32','rbALL_b32','base_b64','rbALL_b64'):
    try:
        d=json.load(open('gpurun_out/r3e_%s.json'%v)); print(v, d['step_us'], round(d['value']))
    except Exception as e: print(v,'ERR',e)"
