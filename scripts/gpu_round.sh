#!/bin/bash
# One gpurun call: GPU tests, the default bench line, the launch list and one ncu --set full capture of the lm_head GEMM.
# usage: scripts/gpu_round.sh <tag>
tag=${1:-r2}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; echo "tests rc=$?" | tee -a gpurun_out/${tag}_tests.log
python bench.py > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/${tag}_bench_ref.json 2> gpurun_out/${tag}_bench_ref.err; echo "ref rc=$?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${tag}_launches.csv \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-sharded --no-gpu-reference --no-full-cycle > gpurun_out/${tag}_ncu_l.log 2>&1; echo "ncu list rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gemm_ -s 60 -c 24 -o gpurun_out/${tag}_gemm -f \
  python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-sharded --no-gpu-reference --no-full-cycle > gpurun_out/${tag}_ncu_g.log 2>&1; echo "ncu full rc=$?"
tail -3 gpurun_out/${tag}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${tag}_smoke.log 2>&1; echo "smoke rc=$?"
DFLASH_LIB=$PWD/build/lib_trace_new.so python scripts/step_trace.py > gpurun_out/${tag}_trace.txt 2>&1; echo "trace rc=$?"
