mkdir -p gpurun_out/r2
timeout 900 python -m pytest tests/test_sdf_build.py -x -q -m gpu -k "tensor_core" -s > gpurun_out/r2/t11_tc.log 2>&1; echo tc rc=$?; grep -E "differing|passed|failed|Error|assert" gpurun_out/r2/t11_tc.log | head -20
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "c1_single or trace_variants or march_work_order or blend" > gpurun_out/r2/t11_quick.log 2>&1; echo quick rc=$?; tail -2 gpurun_out/r2/t11_quick.log
for cfg in "c4 0 a" "c4 0x800 tc" "c5 0 a" "c5 0x800 tc"; do set -- $cfg
timeout 300 python bench.py --workload $1 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --extra-flags $2 > gpurun_out/r2/bench_$1_v26$3.json 2> gpurun_out/r2/bench_$1_v26$3.err; echo $1 $2 rc=$?
python - $1 $3 <<'PY'
import json,sys
w,t=sys.argv[1:3]
try:
    d=json.loads(open(f"gpurun_out/r2/bench_{w}_v26{t}.json").read().strip().splitlines()[-1])
    print(w, t, round(d["ms_per_update"],3), {k:round(v,3) for k,v in d["stage_ms"].items() if isinstance(v,float)})
except Exception as e: print(w, "ERR", e)
PY
done
