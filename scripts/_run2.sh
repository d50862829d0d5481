python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2c_smoke.log 2>&1; echo "smoke rc=$?"
python -m pytest tests/test_gpu_parity.py -x -q -s -k "long_context or full_size_batched" > gpurun_out/r2c_tests.log 2>&1; echo "tests rc=$?"
python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 9000 --csv --log-file gpurun_out/r2c_launches_r64.csv \
  python bench.py --requests 64 --steps 3 --warmup 3 --no-cpu-baseline --no-sharded --no-gpu-reference --no-full-cycle > gpurun_out/r2c_ncu_r64.log 2>&1; echo "ncu r64 rc=$?"
tail -3 gpurun_out/r2c_smoke.log; grep -E "S=|worst|passed|failed" gpurun_out/r2c_tests.log | tail -20
