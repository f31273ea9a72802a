mkdir -p gpurun_out/r2
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 10 --warmup 4 > gpurun_out/r2/bench_c5_v62_n8.json 2> gpurun_out/r2/bench_c5_v62_n8.err; echo bench rc=$?
tail -1 gpurun_out/r2/bench_c5_v62_n8.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2/bench_c5_v62_n8.json").read().strip().splitlines()[-1])
print(round(d["ms_per_update"],2), "e2e", round(d["e2e"]["ms_per_update"],2), d["multi_gpu_parity"]["status"], d["per_rank_trace_blend_ms"], d["parallelism"]["layout"], d["value"])
PY
