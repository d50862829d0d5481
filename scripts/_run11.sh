B="--no-cpu-baseline --no-sharded --no-gpu-reference --no-full-cycle"
python bench.py --requests 64 --steps 30 --warmup 5 $B > gpurun_out/r2n_b64.json 2>gpurun_out/r2n_b64.err
python bench.py --requests 16 --steps 50 --warmup 5 $B > gpurun_out/r2n_b16.json 2>gpurun_out/r2n_b16.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2n_launches_b64.csv \
  python bench.py --requests 64 --steps 4 --warmup 3 $B > gpurun_out/r2n_ncu64.log 2>&1; echo "ncu64 rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2n_launches_b16.csv \
  python bench.py --requests 16 --steps 4 --warmup 3 $B > gpurun_out/r2n_ncu16.log 2>&1; echo "ncu16 rc=$?"
python scripts/parse_launches.py gpurun_out/r2n_launches_b64.csv
python scripts/parse_launches.py gpurun_out/r2n_launches_b16.csv
python -c "
import json
for v in ('b64','b16'):
    d=json.load(open('gpurun_out/r2n_%s.json'%v)); print(v, d['step_us'], d['value'], d['e2e']['value'])"
