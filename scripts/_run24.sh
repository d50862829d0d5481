timeout 900 python -m pytest tests -m gpu -x -q -k "ragged_vs_oracle or full_size_batched or injection_forms" > gpurun_out/r3f_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r3f_tests.log
B="--no-cpu-baseline --no-sharded --no-gpu-reference --no-full-cycle"
for R in 8 16 64; do
python bench.py --requests $R --steps 40 --warmup 5 $B > gpurun_out/r3f_new_b$R.json 2>/dev/null
DFLASH_LIB=$PWD/build/libdflash_rbALL.so python bench.py --requests $R --steps 40 --warmup 5 $B > gpurun_out/r3f_rbALL_b$R.json 2>/dev/null
done
python -c "
import json
for v in ('new_b8','rbALL_b8','new_b16','rbALL_b16','new_b64','rbALL_b64'):
    try:
        d=json.load(open('gpurun_out/r3f_%s.json'%v)); print(v, d['step_us'], round(d['value']))
    except Exception as e: print(v,'ERR',e)"
