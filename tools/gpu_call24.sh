mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2/t27_gpu_all.log 2>&1; echo gpu tests rc=$?; tail -4 gpurun_out/r2/t27_gpu_all.log
( time timeout 900 python bench.py > gpurun_out/r2/bench_c5_v40_default.json 2> gpurun_out/r2/bench_c5_v40_default.err ) 2>&1 | grep real; echo bench rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2/bench_c5_v40_default.json").read().strip().splitlines()[-1])
print(d["config"]["workload"], "ms", round(d["ms_per_update"],2), "value", d["value"], "e2e", d["e2e"]["ms_per_update"] if d["e2e"] else None, "cpu", d.get("cpu_baseline",{}).get("value"), "roof", d["roofline"]["frac"], d["roofline"]["traffic"], "tc", d["blend_tc"]["blend_ms"] if d.get("blend_tc") else None, "launches", d["gpu_launches"], d["clocks"])
print({k:round(v,3) for k,v in d["stage_ms"].items() if isinstance(v,float)})
PY
( time timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2/bench_c5_v40_reference.json 2> gpurun_out/r2/bench_c5_v40_reference.err ) 2>&1 | grep real; tail -c 600 gpurun_out/r2/bench_c5_v40_reference.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
