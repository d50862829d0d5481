python -m pytest tests/test_gpu_parity.py -x -q -k "long_context or full_size_batched or attention_bias or graphed_targets or fresh_engine or batch_lossless" > gpurun_out/r2b_tests.log 2>&1; echo "tests rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2b_smoke.log 2>&1; echo "smoke rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2b_launches_r64.csv \
  python bench.py --requests 64 --steps 3 --warmup 3 --no-cpu-baseline --no-sharded --no-gpu-reference --no-full-cycle > gpurun_out/r2b_ncu_r64.log 2>&1; echo "ncu r64 rc=$?"
tail -5 gpurun_out/r2b_tests.log; tail -3 gpurun_out/r2b_smoke.log
