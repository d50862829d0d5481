python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 --no-gpu-reference --no-full-cycle > gpurun_out/r2y_n2.json 2>gpurun_out/r2y_n2.err; echo "n2 rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2y_n2_ref.json 2>gpurun_out/r2y_n2_ref.err; echo "n2 ref rc=$?"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 scripts/dp_generate_check.py > gpurun_out/r2y_dp.log 2>&1; echo "dp rc=$?"
tail -3 gpurun_out/r2y_dp.log
python -c "
import json
d=json.load(open('gpurun_out/r2y_n2.json')); print({k:d[k] for k in ('value','n_gpus','ms_per_step','step_us','gather','clocks')}); print(d['e2e']); print(d['sharded_batch'])
r=json.load(open('gpurun_out/r2y_n2_ref.json')); print({k:r.get(k) for k in ('value','impl','cpu_baseline')})"
