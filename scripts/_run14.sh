timeout 900 python -m pytest tests -m gpu -x -q -k "ragged_vs_oracle or replays_oracle or spec_generate" > gpurun_out/r2s_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r2s_tests.log
B="--no-cpu-baseline --no-sharded --no-gpu-reference --no-full-cycle"
for i in 1 2; do
python bench.py --steps 200 --warmup 20 $B > gpurun_out/r2s_inj_$i.json 2>gpurun_out/r2s_inj_$i.err
python bench.py --steps 200 --warmup 20 $B --no-inject > gpurun_out/r2s_noinj_$i.json 2>gpurun_out/r2s_noinj_$i.err
done
for R in 16 64; do
python bench.py --requests $R --steps 40 --warmup 5 $B > gpurun_out/r2s_inj_b$R.json 2>gpurun_out/r2s_inj_b$R.err
python bench.py --requests $R --steps 40 --warmup 5 $B --no-inject > gpurun_out/r2s_noinj_b$R.json 2>gpurun_out/r2s_noinj_b$R.err
done
DFLASH_LIB=$PWD/build/lib_trace_new.so python scripts/step_trace.py > gpurun_out/r2s_trace.txt 2>&1
python -c "
import json
for v in ('inj_1','noinj_1','inj_2','noinj_2','inj_b16','noinj_b16','inj_b64','noinj_b64'):
    try:
        d=json.load(open('gpurun_out/r2s_%s.json'%v)); print(v, d['step_us'], round(d['value']), round(d['e2e']['value']), d['launches_per_step'])
    except Exception as e: print(v,'ERR',e)"
head -12 gpurun_out/r2s_trace.txt
