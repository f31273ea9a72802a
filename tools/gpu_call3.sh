mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ragged or c1_single or march_work_order" > gpurun_out/r2/t3_quick.log 2>&1; echo quick rc=$?; tail -3 gpurun_out/r2/t3_quick.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"march_kernel|shade_sorted_kernel|classify_kernel|scatter_kernel" -s 4 -c 4 -o gpurun_out/r2/r2_v20_c5_full -f python bench.py --workload c5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2/ncu_c5_v20_full.log 2>&1; echo ncu rc=$?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"march_kernel" -s 1 -c 1 -o gpurun_out/r2/r2_v20_c4_march -f python bench.py --workload c4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2/ncu_c4_v20_march.log 2>&1; echo ncu rc=$?
ls -la gpurun_out/r2
