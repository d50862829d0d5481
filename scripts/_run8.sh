DFLASH_LIB=$PWD/build/lib_trace_new.so python scripts/step_trace.py > gpurun_out/r2l_trace_new.txt 2>&1
head -24 gpurun_out/r2l_trace_new.txt
