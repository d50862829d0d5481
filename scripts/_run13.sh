timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2q_tests.log 2>&1; echo "tests rc=$?"
tail -5 gpurun_out/r2q_tests.log
B="--no-cpu-baseline --no-sharded --no-gpu-reference --no-full-cycle"
for i in 1 2; do
python bench.py --steps 200 --warmup 20 $B > gpurun_out/r2q_inj_$i.json 2>gpurun_out/r2q_inj_$i.err
python bench.py --steps 200 --warmup 20 $B --no-inject > gpurun_out/r2q_noinj_$i.json 2>gpurun_out/r2q_noinj_$i.err
done
for R in 16 64; do
python bench.py --requests $R --steps 40 --warmup 5 $B > gpurun_out/r2q_inj_b$R.json 2>gpurun_out/r2q_inj_b$R.err
python bench.py --requests $R --steps 40 --warmup 5 $B --no-inject > gpurun_out/r2q_noinj_b$R.json 2>gpurun_out/r2q_noinj_b$R.err
done
DFLASH_ATTN_SPLITS=3 python bench.py --requests 16 --steps 40 --warmup 5 $B > gpurun_out/r2q_inj_b16_s3.json 2>/dev/null
DFLASH_LIB=$PWD/build/lib_trace_new.so python scripts/step_trace.py > gpurun_out/r2q_trace.txt 2>&1
python -c "
import json
for v in ('inj_1','noinj_1','inj_2','noinj_2','inj_b16','noinj_b16','inj_b16_s3','inj_b64','noinj_b64'):
    try:
        d=json.load(open('gpurun_out/r2q_%s.json'%v)); print(v, d['step_us'], round(d['value']), round(d['e2e']['value']), d['launches_per_step'])
    except Exception as e: print(v,'ERR',e)"
