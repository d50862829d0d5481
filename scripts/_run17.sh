timeout 1500 python -m pytest tests -m gpu -x -q -k "injection_forms or ragged_vs_oracle or full_size_batched or spec_generate_batch" > gpurun_out/r2w_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2w_tests.log
B="--no-cpu-baseline --no-sharded --no-gpu-reference --no-full-cycle"
for R in 64 32 16; do
python bench.py --requests $R --steps 40 --warmup 5 $B > gpurun_out/r2w_new_b$R.json 2>gpurun_out/r2w_new_b$R.err
DFLASH_LIB=$PWD/build/libdflash_prev.so python bench.py --requests $R --steps 40 --warmup 5 $B > gpurun_out/r2w_prev_b$R.json 2>gpurun_out/r2w_prev_b$R.err
done
python -c "
import json
for v in ('new_b64','prev_b64','new_b32','prev_b32','new_b16','prev_b16'):
    try:
        d=json.load(open('gpurun_out/r2w_%s.json'%v)); print(v, d['step_us'], round(d['value']), round(d['e2e']['value']), d['launches_per_step'])
    except Exception as e: print(v,'ERR',e)"
