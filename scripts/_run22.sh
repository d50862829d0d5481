B="--no-cpu-baseline --no-sharded --no-gpu-reference --no-full-cycle"
for i in 1 2; do
python bench.py --steps 200 --warmup 20 $B > gpurun_out/r3b_base_$i.json 2>/dev/null
for v in A B C; do
DFLASH_LIB=$PWD/build/lib_trig$v.so python bench.py --steps 200 --warmup 20 $B > gpurun_out/r3b_${v}_$i.json 2>/dev/null
done
done
python -c "
import json
for v in ('base_1','A_1','B_1','C_1','base_2','A_2','B_2','C_2'):
    try:
        d=json.load(open('gpurun_out/r3b_%s.json'%v)); print(v, d['step_us'], round(d['value']))
    except Exception as e: print(v,'ERR',e)"
