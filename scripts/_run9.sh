B="--steps 200 --warmup 20 --no-cpu-baseline --no-sharded --no-gpu-reference --no-full-cycle"
for i in 1 2; do
for v in libdflash_old lib_v1 lib_v2 lib_v3; do
DFLASH_LIB=$PWD/build/$v.so python bench.py $B > gpurun_out/r2j_${v}_$i.json 2>/dev/null
done
python bench.py $B > gpurun_out/r2j_new_$i.json 2>/dev/null
done
python -c "
import json
for v in ('libdflash_old','new','lib_v1','lib_v2','lib_v3'):
    for i in (1,2):
        d=json.load(open('gpurun_out/r2j_%s_%d.json'%(v,i))); print(v,i, d['step_us']['median'], d['e2e']['value'])"
