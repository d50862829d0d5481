python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2d_smoke.log 2>&1; echo "smoke rc=$?"
python scripts/batched_target_profile.py 64 512 > gpurun_out/r2d_bt64.log 2>&1; echo "bt64 rc=$?"
python scripts/batched_target_profile.py 16 512 > gpurun_out/r2d_bt16.log 2>&1; echo "bt16 rc=$?"
tail -2 gpurun_out/r2d_smoke.log; grep "per forward" gpurun_out/r2d_bt64.log gpurun_out/r2d_bt16.log
