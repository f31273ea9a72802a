mkdir -p gpurun_out/r2
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"march_kernel|shade_sorted_kernel|classify_rows_kernel|scatter_kernel|blend_irradiance_lists_kernel|blend_depth_lists_kernel" -s 6 -c 6 -o gpurun_out/r2/r2_v35_c5_full -f python bench.py --workload c5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2/ncu_c5_v35_full.log 2>&1; echo ncu rc=$?
tail -3 gpurun_out/r2/ncu_c5_v35_full.log
ls -la gpurun_out/r2/r2_v35_c5_full.ncu-rep
