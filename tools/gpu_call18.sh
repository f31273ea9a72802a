mkdir -p gpurun_out/r2
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "list_blend or blend or c1_single or ragged or large_probe or full_size" > gpurun_out/r2/t37_blend.log 2>&1; echo blend tests rc=$?; tail -5 gpurun_out/r2/t37_blend.log
for cfg in "c4 0 lists" "c5 0 lists"; do set -- $cfg
timeout 300 python bench.py --workload $1 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --extra-flags $2 > gpurun_out/r2/bench_$1_v53$3.json 2> gpurun_out/r2/bench_$1_v53$3.err; echo $1 $2 rc=$?
python - $1 $3 <<'PY'
import json,sys
w,t=sys.argv[1:3]
try:
    d=json.loads(open(f"gpurun_out/r2/bench_{w}_v53{t}.json").read().strip().splitlines()[-1])
    print(w, t, round(d["ms_per_update"],3), {k:round(v,3) for k,v in d["stage_ms"].items() if isinstance(v,float)})
except Exception as e: print(w, "ERR", e)
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fma.sum --clock-control none -k regex:"blend_" -s 10 -c 10 --csv --log-file gpurun_out/r2/launches_c5_v53_blend.csv python bench.py --workload c5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2/ncu_c5_v53.log 2>&1; echo ncu rc=$?
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2/launches_c5_v53_blend.csv")) if len(r)>10 and r[0].isdigit()]
agg={}
for r in rows:
    k=(int(r[0]), r[4].split("(")[0][-36:]); agg.setdefault(k,{})[r[-3]]=(r[-1])
for k,v in sorted(agg.items()): print(k, {a[:28]:b for a,b in v.items()})
PY
