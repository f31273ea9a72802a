mkdir -p gpurun_out/r2
timeout 240 python -m pytest tests/test_sdf_build.py -x -q -m gpu -k "tensor_core" -s > gpurun_out/r2/t38_umma.log 2>&1; echo umma rc=$?; grep -E "differing|passed|failed|Error" gpurun_out/r2/t23_umma.log | tail -14
for cfg in "c5 0x800 umma" "c4 0x800 umma"; do set -- $cfg
timeout 300 python bench.py --workload $1 --steps 4 --warmup 3 --no-e2e --no-cpu-baseline --extra-flags $2 > gpurun_out/r2/bench_$1_v55$3.json 2> gpurun_out/r2/bench_$1_v55$3.err; echo $1 $2 rc=$?
python - $1 $3 <<'PY'
import json,sys
w,t=sys.argv[1:3]
try:
    d=json.loads(open(f"gpurun_out/r2/bench_{w}_v55{t}.json").read().strip().splitlines()[-1])
    print(w, t, round(d["ms_per_update"],3), {k:round(v,3) for k,v in d["stage_ms"].items() if isinstance(v,float)})
except Exception as e: print(w, "ERR", e)
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_tensor.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"blend_" -s 8 -c 8 --csv --log-file gpurun_out/r2/launches_c5_v55_tc.csv python bench.py --workload c5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --extra-flags 0x800 > gpurun_out/r2/ncu_c5_v55.log 2>&1; echo ncu rc=$?
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2/launches_c5_v55_tc.csv")) if len(r)>10 and r[0].isdigit()]
agg={}
for r in rows:
    k=(int(r[0]), r[4].split("(")[0][-40:]); agg.setdefault(k,{})[r[-3]]=(r[-1])
for k,v in sorted(agg.items()):
    if 'umma' in k[1] or '_tc_' in k[1]: print(k, {a[:30]:b for a,b in v.items()})
PY
timeout 300 python bench.py --workload c5 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-tc-ab > gpurun_out/r2/bench_c5_v55_default.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2/bench_c5_v55_default.json").read().strip().splitlines()[-1])
print("c5 default", round(d["ms_per_update"],3), {k:round(v,3) for k,v in d["stage_ms"].items() if isinstance(v,float)})
PY
