mkdir -p gpurun_out/r2
for sh in blocks8 blocks4; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 8 --steps 10 --warmup 4 --shards $sh --no-e2e > gpurun_out/r2/bench_c5_v57_n8_$sh.json 2> gpurun_out/r2/bench_c5_v57_n8_$sh.err; echo bench $sh rc=$?
tail -1 gpurun_out/r2/bench_c5_v57_n8_$sh.err
python - $sh <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/r2/bench_c5_v57_n8_{sys.argv[1]}.json").read().strip().splitlines()[-1])
print(sys.argv[1], round(d["ms_per_update"],2), d["multi_gpu_parity"]["status"], d["per_rank_trace_blend_ms"], d["parallelism"]["layout"])
PY
done
