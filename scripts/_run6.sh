timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2g_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2g_tests.log
for i in 1 2; do
python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-sharded --no-gpu-reference --no-full-cycle > gpurun_out/r2g_b1_$i.json 2> gpurun_out/r2g_b1_$i.err; echo "b1 rc=$?"
done
python bench.py --requests 8 --steps 100 --warmup 10 --no-cpu-baseline --no-sharded --no-gpu-reference --no-full-cycle > gpurun_out/r2g_b8.json 2> gpurun_out/r2g_b8.err
python bench.py --requests 64 --steps 40 --warmup 5 --no-cpu-baseline --no-sharded --no-gpu-reference --no-full-cycle > gpurun_out/r2g_b64.json 2> gpurun_out/r2g_b64.err
python -c "
import json
for f in ('r2g_b1_1','r2g_b1_2','r2g_b8','r2g_b64'):
    d=json.load(open('gpurun_out/%s.json'%f)); print(f, d['step_us'], d['launches_per_step'], d['e2e']['value'])"
