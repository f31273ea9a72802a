mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "ragged or c1_single or march_work_order or trace_variants or cascaded or full_size or city_with_sky or axis_aligned or z_slab" > gpurun_out/r2/t6_quick.log 2>&1; echo quick rc=$?; tail -3 gpurun_out/r2/t6_quick.log
for cfg in "c4 0 a" "c4 0x200 b" "c4 0x400 c" "c5 0 a" "c5 0x200 b" "c5 0x400 c"; do set -- $cfg
timeout 300 python bench.py --workload $1 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline --extra-flags $2 > gpurun_out/r2/bench_$1_v23$3.json 2> gpurun_out/r2/bench_$1_v23$3.err; echo $1 $2 rc=$?
python - $1 $3 <<'PY'
import json,sys
w,t=sys.argv[1:3]
try:
    d=json.loads(open(f"gpurun_out/r2/bench_{w}_v23{t}.json").read().strip().splitlines()[-1])
    print(w, t, round(d["ms_per_update"],3), {k:round(v,3) for k,v in d["stage_ms"].items() if isinstance(v,float)})
except Exception as e: print(w, "ERR", e)
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,l1tex__t_sector_hit_rate.pct,l1tex__m_l1tex2xbar_req_cycles_active.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"march_kernel|classify_kernel|scatter_kernel|shade_sorted_kernel" -s 4 -c 4 --csv --log-file gpurun_out/r2/launches_c5_v23.csv python bench.py --workload c5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2/ncu_c5_v23.log 2>&1; echo ncu rc=$?
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2/launches_c5_v23.csv")) if len(r)>10 and r[0].isdigit()]
agg={}
for r in rows:
    k=(int(r[0]), r[4].split("(")[0][-36:]); agg.setdefault(k,{})[r[-3]]=(r[-1], r[-2])
for k,v in sorted(agg.items()): print(k, {a[:34]:b for a,b in v.items()})
PY
