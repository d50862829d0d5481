timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2m_tests.log
B="--steps 200 --warmup 20 --no-cpu-baseline --no-sharded --no-gpu-reference --no-full-cycle"
for i in 1 2; do
DFLASH_LIB=$PWD/build/libdflash_old.so python bench.py $B > gpurun_out/r2m_old_$i.json 2>/dev/null
python bench.py $B > gpurun_out/r2m_new_$i.json 2>/dev/null
done
python -c "
import json
for v in ('old_1','new_1','old_2','new_2'):
    d=json.load(open('gpurun_out/r2m_%s.json'%v)); print(v, d['step_us'], d['e2e']['value'])"
