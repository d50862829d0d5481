timeout 1500 python -m pytest tests -m gpu -x -q -k "spec_generate or graphed_target or injection_forms or dflash_generate or candidates" > gpurun_out/r2x_tests.log 2>&1; echo "tests rc=$?"
tail -4 gpurun_out/r2x_tests.log
python bench.py --steps 100 --warmup 10 --no-cpu-baseline --no-sharded --no-gpu-reference > gpurun_out/r2x_bench.json 2>gpurun_out/r2x_bench.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/r2x_bench.json')); print(d['step_us']); print(json.dumps(d['full_cycle'])[:1500])"
