ncu --set full --clock-control none --import-source on -k regex:"finalize_rows_cluster_kernel|qkv_post_kernel|verify_fused_kernel|attn_split_kernel|attn_combine_kernel" -s 700 -c 28 -o gpurun_out/r2f_small64 -f \
  python bench.py --requests 64 --steps 3 --warmup 3 --no-cpu-baseline --no-sharded --no-gpu-reference --no-full-cycle > gpurun_out/r2f_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/r2f_small64.ncu-rep
