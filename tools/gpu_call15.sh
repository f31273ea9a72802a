mkdir -p gpurun_out/r2
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 tests/multi_gpu_check.py > gpurun_out/r2/multi_gpu_check_n8.log 2>&1; echo check rc=$?; grep -E "rank|Error" gpurun_out/r2/multi_gpu_check_n8.log | head -9
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 10 --warmup 4 > gpurun_out/r2/bench_c5_v28_n8.json 2> gpurun_out/r2/bench_c5_v28_n8.err; echo bench rc=$?
tail -2 gpurun_out/r2/bench_c5_v28_n8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2/bench_c5_v28_n8.json").read().strip().splitlines()[-1])
print(round(d["ms_per_update"],2), d["stage_ms"], "e2e", d["e2e"]["ms_per_update"], d["multi_gpu_parity"], d["per_rank_trace_blend_ms"])
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29536 bench.py --gpus 8 --steps 20 --warmup 5 --workload c4 > gpurun_out/r2/bench_c4_v28_n8.json 2> gpurun_out/r2/bench_c4_v28_n8.err; echo c4 rc=$?
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2/bench_c4_v28_n8.json").read().strip().splitlines()[-1])
print(round(d["ms_per_update"],3), d["stage_ms"], "e2e", d["e2e"]["ms_per_update"], d["multi_gpu_parity"])
PY
