mkdir -p gpurun_out/r2
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"blend_.*tc_kernel" -s 2 -c 2 -o gpurun_out/r2/r2_v26_c5_blend_tc -f python bench.py --workload c5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --extra-flags 0x800 > gpurun_out/r2/ncu_c5_v26_tc.log 2>&1; echo ncu rc=$?
