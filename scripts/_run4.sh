for cfg in "64 512" "16 512" "1 512" "1 2304" "64 2304"; do
  python scripts/batched_target_profile.py $cfg > gpurun_out/r2e_bt_${cfg// /_}.log 2>&1; echo "bt $cfg rc=$?"
done
python -m pytest tests/test_gpu_parity.py -x -q -k "graphed_targets or batch_lossless" > gpurun_out/r2e_tests.log 2>&1; echo "tests rc=$?"
grep -h "per forward" gpurun_out/r2e_bt_*.log; tail -3 gpurun_out/r2e_tests.log
