mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_sdf_build.py -x -q -m gpu > gpurun_out/r2/t16_parity.log 2>&1; echo parity rc=$?; tail -3 gpurun_out/r2/t16_parity.log
for cfg in "c4 0 a" "c5 0 a"; do set -- $cfg
timeout 300 python bench.py --workload $1 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --extra-flags $2 > gpurun_out/r2/bench_$1_v29$3.json 2> gpurun_out/r2/bench_$1_v29$3.err; echo $1 $2 rc=$?
python - $1 $3 <<'PY'
import json,sys
w,t=sys.argv[1:3]
try:
    d=json.loads(open(f"gpurun_out/r2/bench_{w}_v29{t}.json").read().strip().splitlines()[-1])
    print(w, t, round(d["ms_per_update"],3), {k:round(v,3) for k,v in d["stage_ms"].items() if isinstance(v,float)})
except Exception as e: print(w, "ERR", e)
PY
done
