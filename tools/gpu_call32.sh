mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2/t41_gpu_all.log 2>&1; echo gpu tests rc=$?; tail -3 gpurun_out/r2/t41_gpu_all.log
( time timeout 900 python bench.py > gpurun_out/r2/bench_c5_v61_default.json 2> gpurun_out/r2/bench_c5_v61_default.err ) 2>&1 | grep real; echo bench rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2/bench_c5_v61_default.json").read().strip().splitlines()[-1])
print(d["config"]["workload"], "ms", round(d["ms_per_update"],2), "value", d["value"], "e2e", d["e2e"]["ms_per_update"] if d["e2e"] else None, "cpu", d.get("cpu_baseline",{}).get("value"), "roof", d["roofline"]["frac"], d["roofline"]["traffic"], "tc", d["blend_tc"], "launches", d["gpu_launches"], d["clocks"])
print({k:round(v,3) for k,v in d["stage_ms"].items() if isinstance(v,float)})
PY
timeout 300 python bench.py --workload c4 > gpurun_out/r2/bench_c4_v61_default.json 2> gpurun_out/r2/bench_c4_v61_default.err; python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2/bench_c4_v61_default.json").read().strip().splitlines()[-1])
print("c4 ms", round(d["ms_per_update"],3), "e2e", d["e2e"]["ms_per_update"], {k:round(v,3) for k,v in d["stage_ms"].items() if isinstance(v,float)}, d["blend_tc"]["blend_ms"])
PY
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2/bench_c5_v61_reference.json 2> gpurun_out/r2/bench_c5_v61_reference.err ) 2>&1 | grep real
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"march_kernel|shade_sorted_kernel|classify_rows_kernel|scatter_kernel|blend_irradiance_lists_kernel|blend_depth_lists_kernel" -s 6 -c 6 -o gpurun_out/r2/r2_v61_c5_full -f python bench.py --workload c5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-tc-ab > gpurun_out/r2/ncu_c5_v61_full.log 2>&1; echo ncu rc=$?
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio --clock-control none -k regex:"march_kernel|classify|scan_|scatter_kernel|shade_sorted_kernel|blend_" -s 15 -c 15 --csv --log-file gpurun_out/r2/launches_c5_v61.csv python bench.py --workload c5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-tc-ab > gpurun_out/r2/ncu_c5_v61.log 2>&1; echo ncu list rc=$?
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
