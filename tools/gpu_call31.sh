mkdir -p gpurun_out/r2
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "list_blend_equals_oracle_and_tiled_blend and 50 or test_ragged" > gpurun_out/r2/racecheck_list_blend.log 2>&1; echo racecheck rc=$?; grep -E "RACECHECK SUMMARY|passed|failed" gpurun_out/r2/racecheck_list_blend.log | head -6; grep -E "Race reported" gpurun_out/r2/racecheck_list_blend.log | sort | uniq -c | head
timeout 300 python -m pytest tests/test_gpu_parity.py -x -q -m gpu -k "trace_variants or work_order or full_size" 2>&1 | tail -2
timeout 300 python bench.py --workload c5 --steps 5 --warmup 3 --no-e2e --no-cpu-baseline --no-tc-ab 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('c5', round(d['ms_per_update'],3), {k:round(v,3) for k,v in d['stage_ms'].items() if isinstance(v,float)})"
