timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2k_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2k_tests.log
B="--steps 200 --warmup 20 --no-cpu-baseline --no-sharded --no-gpu-reference --no-full-cycle"
for i in 1 2; do
DFLASH_LIB=$PWD/build/libdflash_old.so python bench.py $B > gpurun_out/r2k_old_$i.json 2>/dev/null
python bench.py $B > gpurun_out/r2k_new_$i.json 2>/dev/null
done
DFLASH_LIB=$PWD/build/libdflash_old.so python bench.py --requests 64 --steps 40 --warmup 5 --no-cpu-baseline --no-sharded --no-gpu-reference --no-full-cycle > gpurun_out/r2k_old_b64.json 2>/dev/null
python bench.py --requests 64 --steps 40 --warmup 5 --no-cpu-baseline --no-sharded --no-gpu-reference --no-full-cycle > gpurun_out/r2k_new_b64.json 2>/dev/null
python -c "
import json
for v in ('old_1','new_1','old_2','new_2','old_b64','new_b64'):
    d=json.load(open('gpurun_out/r2k_%s.json'%v)); print(v, d['step_us'], d['e2e']['value'])"
