mkdir -p gpurun_out/r2
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2/t10_gpu_all.log 2>&1; echo gpu-tests rc=$?; tail -5 gpurun_out/r2/t10_gpu_all.log
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all --print-limit 20 python __graft_entry__.py --smoke > gpurun_out/r2/racecheck_smoke.log 2>&1; echo racecheck rc=$?; tail -6 gpurun_out/r2/racecheck_smoke.log
timeout 600 python bench.py > gpurun_out/r2/bench_c5_v25.json 2> gpurun_out/r2/bench_c5_v25.err; echo bench rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2/bench_c5_v25.json").read().strip().splitlines()[-1])
print(round(d["ms_per_update"],2), {k:round(v,2) for k,v in d["stage_ms"].items() if isinstance(v,float)}, "e2e", round(d["e2e"]["ms_per_update"],2), "cpu", d["cpu_baseline"]["value"], d["cpu_baseline"]["counters_per_ray"])
PY
