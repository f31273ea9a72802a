mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_sdf_build.py -x -q -m gpu -k "trace_variants or ragged or work_order or full_size or c1_ or tensor_core" > gpurun_out/r2/t40_trace.log 2>&1; echo trace tests rc=$?; tail -3 gpurun_out/r2/t40_trace.log
for cfg in "c4 0 a" "c5 0 a"; do set -- $cfg
timeout 300 python bench.py --workload $1 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --extra-flags $2 > gpurun_out/r2/bench_$1_v59$3.json 2> gpurun_out/r2/bench_$1_v59$3.err; echo $1 $2 rc=$?
python - $1 $3 <<'PY'
import json,sys
w,t=sys.argv[1:3]
try:
    d=json.loads(open(f"gpurun_out/r2/bench_{w}_v59{t}.json").read().strip().splitlines()[-1])
    print(w, t, round(d["ms_per_update"],3), {k:round(v,3) for k,v in d["stage_ms"].items() if isinstance(v,float)})
except Exception as e: print(w, "ERR", e)
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,l1tex__t_sector_hit_rate.pct,lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"shade_sorted" -s 1 -c 1 --csv --log-file gpurun_out/r2/launches_c5_v59_shade.csv python bench.py --workload c5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2/ncu_c5_v59.log 2>&1; echo ncu rc=$?
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2/launches_c5_v59_shade.csv")) if len(r)>10 and r[0].isdigit()]
for r in rows: print(r[4][:30], r[-3], r[-1])
PY
