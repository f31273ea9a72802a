mkdir -p gpurun_out/r2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"march_kernel" -s 1 -c 1 -o gpurun_out/r2/r2_v23_c5_march -f python bench.py --workload c5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --extra-flags 0x200 > gpurun_out/r2/ncu_c5_v23_full.log 2>&1; echo ncu rc=$?
ls -la gpurun_out/r2/*.ncu-rep
