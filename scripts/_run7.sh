B="--steps 200 --warmup 20 --no-cpu-baseline --no-sharded --no-gpu-reference --no-full-cycle"
for i in 1 2 3; do
DFLASH_LIB=$PWD/build/libdflash_old.so python bench.py $B > gpurun_out/r2h_old_$i.json 2>/dev/null
python bench.py $B > gpurun_out/r2h_new_$i.json 2>/dev/null
done
python -c "
import json
for f in ('old_1','new_1','old_2','new_2','old_3','new_3'):
    d=json.load(open('gpurun_out/r2h_%s.json'%f)); print(f, d['step_us'], d['e2e']['value'])"
