mkdir -p gpurun_out/r2
timeout 600 python bench.py > gpurun_out/r2/bench_c5_v64_default.json 2> gpurun_out/r2/bench_c5_v64_default.err; echo rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2/bench_c5_v64_default.json").read().strip().splitlines()[-1])
print(round(d["ms_per_update"],2), d["roofline_issue"], d["roofline"]["frac"], d["e2e"]["ms_per_update"])
PY
timeout 200 python bench.py --workload c4 --no-cpu-baseline --no-tc-ab 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c4', round(d['ms_per_update'],3), d['roofline_issue'])"
