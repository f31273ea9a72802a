mkdir -p gpurun_out/r2
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "c1_single or trace_variants or march_work_order or cascaded or ragged or city_with_sky or axis_aligned" > gpurun_out/r2/t2_quick.log 2>&1; echo quick rc=$?; tail -5 gpurun_out/r2/t2_quick.log
timeout 300 python bench.py --workload c4 --steps 8 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2/bench_c4_v20.json 2> gpurun_out/r2/bench_c4_v20.err; echo c4 rc=$?
timeout 300 python bench.py --workload c5 --steps 6 --warmup 3 --no-e2e --no-cpu-baseline > gpurun_out/r2/bench_c5_v20.json 2> gpurun_out/r2/bench_c5_v20.err; echo c5 rc=$?
python - <<'PY'
import json
for w in ("c4","c5"):
    try:
        d=json.loads(open(f"gpurun_out/r2/bench_{w}_v20.json").read().strip().splitlines()[-1])
        print(w, d["ms_per_update"], d["stage_ms"])
    except Exception as e: print(w, "ERR", e)
PY
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s 13 -c 14 --csv --log-file gpurun_out/r2/launches_c5_v20.csv python bench.py --workload c5 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline > gpurun_out/r2/ncu_c5_v20.log 2>&1; echo ncu rc=$?
python - <<'PY'
import csv
rows=[r for r in csv.reader(open("gpurun_out/r2/launches_c5_v20.csv")) if len(r)>10 and r[0].isdigit()]
agg={}
for r in rows:
    k=(r[0], r[4].split("(")[0][-40:]); agg.setdefault(k,{})[r[-3]]=float(r[-1].replace(",",""))
for k,v in agg.items(): print(k, {a:(b/1e9 if 'bytes' in a else b/1e6) for a,b in v.items()})
PY
