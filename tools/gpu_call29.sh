mkdir -p gpurun_out/r2
for v in shade4 shade6; do
  if [ $v = default ]; then unset LUX_DDGI_LIB; else export LUX_DDGI_LIB=$PWD/luxgi_b200/variants/lib_$v.so; fi
  for w in c5; do
    timeout 300 python bench.py --workload $w --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-tc-ab > gpurun_out/r2/bench_${w}_v63_$v.json 2> gpurun_out/r2/bench_${w}_v63_$v.err; 
    python - $w $v <<'PY'
import json,sys
w,t=sys.argv[1:3]
try:
    d=json.loads(open(f"gpurun_out/r2/bench_{w}_v63_{t}.json").read().strip().splitlines()[-1])
    print(w, t, round(d["ms_per_update"],3), {k:round(v,3) for k,v in d["stage_ms"].items() if isinstance(v,float)})
except Exception as e: print(w, "ERR", e)
PY
  done
done
