mkdir -p gpurun_out/r2
for cfg in "c5 0x200 u" "c4 0x200 u"; do set -- $cfg
timeout 300 python bench.py --workload $1 --steps 4 --warmup 2 --no-e2e --no-cpu-baseline --extra-flags $2 --unsorted > gpurun_out/r2/bench_$1_v23$3.json 2> gpurun_out/r2/bench_$1_v23$3.err; echo $1 $2 rc=$?
python - $1 $3 <<'PY'
import json,sys
w,t=sys.argv[1:3]
try:
    d=json.loads(open(f"gpurun_out/r2/bench_{w}_v23{t}.json").read().strip().splitlines()[-1])
    print(w, t, round(d["ms_per_update"],3), {k:round(v,3) for k,v in d["stage_ms"].items() if isinstance(v,float)})
except Exception as e: print(w, "ERR", e)
PY
done
