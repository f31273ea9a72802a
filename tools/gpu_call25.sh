mkdir -p gpurun_out/r2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py > gpurun_out/r2/multi_gpu_check_n2_v40.log 2>&1; echo check rc=$?; grep -E "rank|Error" gpurun_out/r2/multi_gpu_check_n2_v40.log | head -4
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 4 > gpurun_out/r2/bench_c5_v40_n2.json 2> gpurun_out/r2/bench_c5_v40_n2.err; echo bench rc=$?
tail -2 gpurun_out/r2/bench_c5_v40_n2.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2/bench_c5_v40_n2.json").read().strip().splitlines()[-1])
print(round(d["ms_per_update"],2), {k:(round(v,3) if isinstance(v,float) else v) for k,v in d["stage_ms"].items() if k!="note"}, "e2e", d["e2e"]["ms_per_update"], d["multi_gpu_parity"], d["per_rank_trace_blend_ms"])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/r2/bench_c5_v40_n2_reference.json 2> gpurun_out/r2/bench_c5_v40_n2_reference.err; echo ref rc=$?; tail -c 300 gpurun_out/r2/bench_c5_v40_n2_reference.json
