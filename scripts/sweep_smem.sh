#!/bin/bash
# Step time vs the TMA pipeline budget of the chained (partials) GEMMs. Variants are built with
#   nvcc ... -DDFLASH_GEMM_SMEM_KB_PARTIALS=<kb> -o build/lib_smem<kb>.so   
for kb in default "$@"; do
  if [ "$kb" = default ]; then unset DFLASH_LIB; else export DFLASH_LIB=$PWD/build/lib_smem$kb.so; fi
  python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-full-cycle > gpurun_out/sweep_smem_$kb.json 2> gpurun_out/sweep_smem_$kb.err
  python -c "
import json;d=json.load(open('gpurun_out/sweep_smem_$kb.json'));print('$kb', d['step_us'], d['roofline']['launch_us'])"
done
