timeout 900 python -m pytest tests -m gpu -x -q -k "replays_oracle or ragged_vs_oracle or full_size_qwen3" > gpurun_out/r2z_tests.log 2>&1; echo "tests rc=$?"
tail -3 gpurun_out/r2z_tests.log
B="--no-cpu-baseline --no-sharded --no-gpu-reference --no-full-cycle"
for i in 1 2; do
python bench.py --steps 200 --warmup 20 $B > gpurun_out/r2z_new_$i.json 2>gpurun_out/r2z_new_$i.err
DFLASH_LIB=$PWD/build/libdflash_prev.so python bench.py --steps 200 --warmup 20 $B > gpurun_out/r2z_prev_$i.json 2>gpurun_out/r2z_prev_$i.err
done
python bench.py --requests 2 --steps 100 --warmup 10 $B > gpurun_out/r2z_new_b2.json 2>/dev/null
DFLASH_LIB=$PWD/build/libdflash_prev.so python bench.py --requests 2 --steps 100 --warmup 10 $B > gpurun_out/r2z_prev_b2.json 2>/dev/null
DFLASH_LIB=$PWD/build/lib_trace_new.so python scripts/step_trace.py > gpurun_out/r2z_trace.txt 2>&1
python -c "
import json
for v in ('new_1','prev_1','new_2','prev_2','new_b2','prev_b2'):
    try:
        d=json.load(open('gpurun_out/r2z_%s.json'%v)); print(v, d['step_us'], round(d['value']), round(d['e2e']['value']), d['launches_per_step'])
    except Exception as e: print(v,'ERR',e)"
sed -n 28,50p gpurun_out/r2z_trace.txt
