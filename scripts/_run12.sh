B="--no-cpu-baseline --no-sharded --no-gpu-reference --no-full-cycle"
ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 1500 --csv --log-file gpurun_out/r3d_launches_b64.csv \
  python bench.py --requests 64 --steps 4 --warmup 3 $B > gpurun_out/r3d_ncu64.log 2>&1; echo "ncu64 rc=$?"
python scripts/parse_launches.py gpurun_out/r3d_launches_b64.csv
